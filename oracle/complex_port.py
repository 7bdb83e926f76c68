"""TEST INFRASTRUCTURE -- CPU restatement of the COMPLEX-isometry extension of the hot path.

The reference's Stiefel algebra is real ("TODO: add back the conj", opentn/stiefel.py:29,36,44,184,201,214,218,320); its
representation layer is already complex (`kraus2superop`, transformations.py:170-180: sum_k E_k (x) conj(E_k);
`unfactorize_psd`, :804-806: x @ x.conj().T).  This module puts the conjugates back, function by function:

    ortho2super      transformations.py:743-749 with x complex:  Y[(a,b),(c,d)] = sum_k x[(a,k),c] conj(x[(b,k),d])
    model / cost     optimization.py:147-151, :56-61 (complex model minus complex target, Frobenius norm)
    egrad            the gradient in the sense df = Re tr(Z^H dx)  (= d/dRe x + i d/dIm x)
    rgrad map        stiefel.py:398-409 with ^T -> ^H:   Z/a0 + (1/a1 - 2/a0)/2 X X^H Z - 1/(2 a1) X Z^H X
    connection       stiefel.py:468-475 with ^T -> ^H
    project          stiefel.py:23-30 with sym -> Hermitian part
    retraction       stiefel.py:323-330 (scipy.linalg.polar handles complex)
    canonical metric stiefel.py:76-83:  Re tr(d1^H (I - X X^H / 2) d2)

Parity is pinned in tests/test_oracle_complex.py: the closed forms below against torch autodiff (complex128 model, real leaf
tensors for Re x and Im x; torch.func.jvp of the gradient for the Hessian-vector product) and against the real oracle
(oracle/port.py) on real inputs.  There is no reference implementation to run: "parity pinned by independent autodiff".
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg


def herm(A):
    return 0.5 * (A + A.conj().T)


def ortho2super(x):
    """sum_k E_k (x) conj(E_k), E_k = x[k::K, :]  (row a*K + k)."""
    n, dim = x.shape
    K = n // dim
    E = x.reshape(dim, K, dim)                       # [a, k, c]
    Y = np.einsum("akc,bkd->abcd", E, E.conj())
    return Y.reshape(dim * dim, dim * dim)


def ortho2super_jvp(x, eta):
    n, dim = x.shape
    K = n // dim
    E, Ed = x.reshape(dim, K, dim), eta.reshape(dim, K, dim)
    Y = np.einsum("akc,bkd->abcd", Ed, E.conj()) + np.einsum("akc,bkd->abcd", E, Ed.conj())
    return Y.reshape(dim * dim, dim * dim)


def ortho2super_vjp(x, W):
    """Z with  Re tr(W^H dY) = Re tr(Z^H dx):  Z[(a,k),c] = sum_{b,d} (W[(a,b),(c,d)] + conj(W[(b,a),(d,c)])) x[(b,k),d]."""
    n, dim = x.shape
    K = n // dim
    E = x.reshape(dim, K, dim)
    W4 = W.reshape(dim, dim, dim, dim)               # [a, b, c, d]
    Ws = W4 + W4.conj().transpose(1, 0, 3, 2)
    return np.einsum("abcd,bkd->akc", Ws, E).reshape(n, dim)


def model(xs):
    M = ortho2super(xs[0])
    for x in xs[1:]:
        M = ortho2super(x) @ M
    return M


def cost(T, xs):
    return np.linalg.norm(model(xs) - T)


def egrad_jvp(T, xs, etas=None):
    """(f, Z) or, with directions, (f, Z, Zdot): closed-form reverse / forward-over-reverse sweeps of the complex chain."""
    m = len(xs)
    Y = [ortho2super(x) for x in xs]
    P = [Y[0]]
    for l in range(1, m):
        P.append(Y[l] @ P[-1])
    R = P[-1] - T
    f = np.linalg.norm(R)
    G = R / f
    tangent = etas is not None
    if tangent:
        Yd = [ortho2super_jvp(x, e) for x, e in zip(xs, etas)]
        Pd = [Yd[0]]
        for l in range(1, m):
            Pd.append(Yd[l] @ P[l - 1] + Y[l] @ Pd[-1])
        fd = np.real(np.vdot(R, Pd[-1])) / f
        Gd = Pd[-1] / f - R * (fd / f ** 2)
    Z, Zd = [None] * m, [None] * m
    for l in range(m - 1, -1, -1):
        W = G @ P[l - 1].conj().T if l > 0 else G
        Z[l] = ortho2super_vjp(xs[l], W)
        if tangent:
            Wd = (Gd @ P[l - 1].conj().T + G @ Pd[l - 1].conj().T) if l > 0 else Gd
            Zd[l] = ortho2super_vjp(xs[l], Wd) + ortho2super_vjp(etas[l], W)
        if l > 0:
            if tangent:
                Gd = Yd[l].conj().T @ G + Y[l].conj().T @ Gd
            G = Y[l].conj().T @ G
    return (f, Z, Zd) if tangent else (f, Z)


def project(X, Z):
    return Z - X @ herm(X.conj().T @ Z)


def gradient_ambient2riemannian(x, g, a0=1.0, a1=1.0):
    return g / a0 + 0.5 * (1 / a1 - 2 / a0) * x @ (x.conj().T @ g) - 0.5 / a1 * x @ (g.conj().T @ x)


def riemannian_connection(D_nu, nu, eta, x, a0=1.0, a1=1.0):
    In = np.eye(x.shape[0])
    return (D_nu + 0.5 * x @ (eta.conj().T @ nu + nu.conj().T @ eta)
            + ((a0 - a1) / a0) * (In - x @ x.conj().T) @ (eta @ nu.conj().T + nu @ eta.conj().T) @ x)


def metric_alphas(metric):
    if metric == "euclidean":
        return 1.0, 1.0
    if metric == "canonical":
        return 1.0, 0.5
    return float(metric[0]), float(metric[1])


def triple(T, xs, etas, metric="canonical"):
    """(f, rgrad, H eta) as (n,p) complex tangent matrices."""
    a0, a1 = metric_alphas(metric)
    c1, c2 = 0.5 * (1 / a1 - 2 / a0), 0.5 / a1
    f, Z, Zd = egrad_jvp(T, xs, etas)
    g, h = [], []
    for x, e, z, zd in zip(xs, etas, Z, Zd):
        xh = x.conj().T
        nu = gradient_ambient2riemannian(x, z, a0, a1)
        dnu = (zd / a0 + c1 * (e @ (xh @ z) + x @ (e.conj().T @ z) + x @ (xh @ zd))
               - c2 * (e @ (z.conj().T @ x) + x @ (zd.conj().T @ x) + x @ (z.conj().T @ e)))
        g.append(nu)
        h.append(project(x, riemannian_connection(dnu, nu, e, x, a0, a1)))
    return f, g, h


def cost_rgrad(T, xs, metric="canonical"):
    a0, a1 = metric_alphas(metric)
    f, Z = egrad_jvp(T, xs)
    return f, [gradient_ambient2riemannian(x, z, a0, a1) for x, z in zip(xs, Z)]


def canonical_dot(xs, a, b):
    tot = 0.0
    for x, u, v in zip(xs, a, b):
        tot += np.real(np.vdot(u, v)) - 0.5 * np.real(np.vdot(x.conj().T @ u, x.conj().T @ v))
    return tot


def retract(xs, etas):
    return [scipy.linalg.polar(x + e)[0] for x, e in zip(xs, etas)]


def random_stiefel(n, p, rng):
    q, r = np.linalg.qr(rng.standard_normal((n, p)) + 1j * rng.standard_normal((n, p)))
    d = np.diag(r)
    return q * (d / np.abs(d))


# ---- coefficient vectors (complex analogue of stiefel.py:266-310): Euclidean dot == canonical metric -----------------------
def get_orthogonal_complement(x):
    n, p = x.shape
    q, _ = np.linalg.qr(x, mode="complete")
    return q[:, p:]


def parametrization_from_tangent(x, z, x_comp=None):
    """[Re A_ij, Im A_ij (i<j), Im A_ii / sqrt2, Re B, Im B] with z = x A + x_perp B, A skew-Hermitian (p^2 real dof)."""
    p = x.shape[1]
    A = x.conj().T @ z
    A = 0.5 * (A - A.conj().T)
    iu = np.triu_indices(p, k=1)
    vec = [A[iu].real, A[iu].imag, np.diag(A).imag / np.sqrt(2.0)]
    if x.shape[0] > p:
        xc = get_orthogonal_complement(x) if x_comp is None else x_comp
        B = xc.conj().T @ z
        vec += [B.real.ravel(), B.imag.ravel()]
    return np.concatenate(vec)


def tangent_from_parametrization(x, vec, x_comp=None):
    n, p = x.shape
    nu = p * (p - 1) // 2
    iu = np.triu_indices(p, k=1)
    A = np.zeros((p, p), dtype=complex)
    A[iu] = vec[:nu] + 1j * vec[nu:2 * nu]
    A = A - A.conj().T
    A[np.diag_indices(p)] = 1j * np.sqrt(2.0) * vec[2 * nu:2 * nu + p]
    z = x @ A
    if n > p:
        xc = get_orthogonal_complement(x) if x_comp is None else x_comp
        rest = vec[2 * nu + p:]
        half = (n - p) * p
        z = z + xc @ (rest[:half] + 1j * rest[half:]).reshape(n - p, p)
    return z
