"""TEST INFRASTRUCTURE ONLY -- CPU (NumPy/SciPy, float64) restatement of the reference's hot path.

This module is the *checker* for the CUDA path in `opentn_b200/`.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import it.
The product package never does (and fails loudly when `libotn.so` is missing).

What it restates (citations are into `/root/reference/`):

* representation layer  -- `opentn/transformations.py:716-749` (ortho2choi / unfactorize_psd / choi2super /
  ortho2super), `:170-180` (kraus2superop), `:405-490` (kron embedding + axis permutation + pbc shift),
  `:841-851` (composition), and the problem set-up `:24-66, 236-402` (Liouvillians, expm, Trotter layers)
  and `:68-94, 701-802` (super2choi, truncated PSD factor, choi2ortho, super2ortho);
* model / cost layer    -- `opentn/optimization.py:56-61, 106-151, 172-185`;
* manifold layer        -- `opentn/stiefel.py:23-45, 76-91, 175-355, 398-544`;
* optimiser layer       -- `opentn/trust_region_rcopt.py:5-166`.

The reference obtains derivatives with `jax.grad` (`stiefel.py:423,437,450`) and `jax.jvp`
(`stiefel.py:520`); jax/jaxlib are a third-party dependency that is absent here (version unpinned by the
reference: it ships no requirements file).  Those three call sites are restated below as closed-form
reverse / forward-over-reverse sweeps of the same computational graph (`egrad`, `egrad_jvp`).

Pinning (see `tests/test_oracle_*.py`):
  * every cost anchor printed in the reference's notebooks / stored in its `.npy` artefacts
    (`tests/golden/reference_goldens.npz`, made by `tests/golden/make_golden.py`),
  * the reference's own code executed under `oracle/ref_shim.py` in the build container
    (forward model, Stiefel algebra, retraction, tCG, the whole TR loop -- bitwise or to 1e-13),
  * `torch.func.jvp(torch.func.grad(.))` of a transcription of the reference cost for `egrad` / `egrad_jvp`,
    and central finite differences.
"""
from __future__ import annotations

import warnings
from itertools import chain
from math import log

import numpy as np
import scipy.linalg

# --------------------------------------------------------------------------------------------------
# constants (opentn/states/qubits.py:5-17)
# --------------------------------------------------------------------------------------------------
I2 = np.eye(2, dtype=complex)
X2 = np.array([[0, 1], [1, 0]], dtype=complex)
Z2 = np.array([[1, 0], [0, -1]], dtype=complex)
Y2 = np.array([[0, -1j], [1j, 0]], dtype=complex)


def get_ladder_operator(num_levels: int = 2, adjoint: bool = False):
    """opentn/states/qubits.py:19-37"""
    a = np.diag(np.array([np.sqrt(i) for i in range(1, num_levels)], dtype=complex), k=1)
    return a.conj().T if adjoint else a


# --------------------------------------------------------------------------------------------------
# problem set-up: Liouvillians and targets (transformations.py:24-66, 236-402)
# --------------------------------------------------------------------------------------------------
def exp_operator_dt(op, tau=1.0):
    """transformations.py:24-32 (both backends are scipy's Pade expm under the shim)."""
    return scipy.linalg.expm(op * tau)


def vectorize_dissipative(L):
    """transformations.py:243-250, row-major vectorisation: L rho L^+ -> (L (x) L*) |rho>>."""
    dim = L.shape[0]
    Id = np.eye(dim, dtype=complex)
    return np.kron(L, L.conj()) - 0.5 * np.kron(L.T.conj() @ L, Id) - 0.5 * np.kron(Id, L.T @ L.conj())


def lindbladian2super(Li):
    """transformations.py:34-66 with H=None."""
    dim = Li[0].shape[0]
    out = np.zeros((dim * dim, dim * dim), dtype=complex)
    for L in Li:
        out += vectorize_dissipative(L)
    return out


def op2fullspace(op, i, N, num_sites=1):
    """transformations.py:252-260"""
    if N == 1:
        return op
    return np.kron(np.eye(2 ** i), np.kron(op, np.eye(2 ** (N - i - num_sites))))


def permute_operator(op, permutation, d):
    """transformations.py:455-464"""
    N = len(permutation)
    t = np.reshape(op, [d] * 2 * N)
    t = np.transpose(t, list(permutation) + [N + p for p in permutation])
    return np.reshape(t, (d ** N, d ** N))


def create_pbc_liouvillian(Lnn, N, d):
    """transformations.py:364-369: the pair acting on (N-2, N-1) is rotated right to act on (N-1, 0)."""
    full = op2fullspace(Lnn, i=N - 2, N=N, num_sites=2)
    perm = list(range(N))
    perm = perm[-1:] + perm[:-1]  # permute_cyclic(..., direction='right'), transformations.py:446-453
    return vectorize_dissipative(permute_operator(full, perm, d))


def local_lindbladian_to_full_liouvillians(Lnn, N, d, pbc=False):
    """transformations.py:337-362"""
    num_sites = int(log(Lnn.shape[0], d))
    if N == num_sites or N == 1:
        pbc = False
    D = d ** (2 * N)
    odd = np.zeros((D, D), dtype=complex)
    for i in range(0, N, 2):
        odd += vectorize_dissipative(op2fullspace(Lnn, i, N, num_sites))
    even = np.zeros((D, D), dtype=complex)
    for i in range(1, N - 1, 2):
        even += vectorize_dissipative(op2fullspace(Lnn, i, N, num_sites))
    if pbc:
        even += create_pbc_liouvillian(Lnn, N, d)
    return even + odd, odd, even


def create_2local_liouvillians(Li, N, d, pbc=False):
    """transformations.py:324-335"""
    if not isinstance(Li, list):
        Li = [Li]
    D = d ** (2 * N)
    odd = np.zeros((D, D), dtype=complex)
    even = np.zeros((D, D), dtype=complex)
    for L in Li:
        _, o, e = local_lindbladian_to_full_liouvillians(L, N, d, pbc)
        odd += o
        even += e
    return odd + even, odd, even


def get_kitaev_nn_linbladian(gamma):
    """transformations.py:371-377"""
    lo, ra = get_ladder_operator(), get_ladder_operator(adjoint=True)
    return np.sqrt(gamma) * (op2fullspace(ra, 0, 2) + op2fullspace(ra, 1, 2)) @ (
        op2fullspace(lo, 0, 2) - op2fullspace(lo, 1, 2)) / 4


def pspl_lindbladians(gamma=1.0):
    """The 'PSPL' Pauli-noise pair of experiments/trust_region_pauli_noise.ipynb cell 11."""
    g = np.sqrt(gamma)
    return [g * (np.kron(X2, I2) - np.kron(I2, X2)), g * (np.kron(Z2, I2) - np.kron(I2, Z2))]


def create_trotter_layers(liouvillians, tau=1.0, half_idx=1):
    """transformations.py:387-402 (order 2)"""
    return [exp_operator_dt(op, tau / 2 if i == half_idx else tau) for i, op in enumerate(liouvillians)]


def get_general_trotter_local_ansatz(lindbladians, tau, n):
    """optimization.py:172-185: [half] + [full]*(2n-1) + [half] local 16x16 superoperators."""
    if not isinstance(lindbladians, list):
        lindbladians = [lindbladians]
    dt = tau / n
    s = lindbladian2super(lindbladians)
    half, full = exp_operator_dt(s, dt / 2), exp_operator_dt(s, dt)
    return [half] + [full] * (2 * n - 1) + [half]


# --------------------------------------------------------------------------------------------------
# representation conversions (transformations.py:68-125, 701-806)
# --------------------------------------------------------------------------------------------------
def super2choi(superop, dim=None):
    """transformations.py:68-94"""
    if not dim:
        dim = int(np.sqrt(superop.shape[0]))
    return np.reshape(superop, [dim] * 4).swapaxes(1, 2).reshape(dim * dim, dim * dim)


choi2super = super2choi  # transformations.py:96-125 is the same involution


def split_matrix_svd(op, chi_max=2, eps=1e-9):
    """transformations.py:752-790"""
    u, s, v = np.linalg.svd(op, full_matrices=False)
    keep = chi_max if eps is None else min(chi_max, int(np.sum(s > eps)))
    idx = np.argsort(s)[::-1][:keep]
    return u[:, idx], s[idx], v[idx, :]


def factorize_psd_truncated(psd, chi_max=2, eps=None):
    """transformations.py:793-802"""
    u, s, _ = split_matrix_svd(psd, chi_max, eps)
    return u @ np.diag(np.sqrt(s))


def choi2ortho(x, dim=None):
    """transformations.py:701-714"""
    if not dim:
        dim = int(np.sqrt(x.shape[0]))
    k = x.shape[1]
    return np.reshape(x, [dim, dim, k]).swapaxes(1, 2).reshape(dim * k, dim)


def ortho2choi(x, dim=None):
    """transformations.py:716-726"""
    if not dim:
        dim = x.shape[1]
    k = x.shape[0] // dim
    return np.reshape(x, [dim, k, dim]).swapaxes(1, 2).reshape(dim * dim, k)


def super2ortho(x, rank=None):
    """transformations.py:728-741"""
    c = super2choi(x)
    if not rank:
        rank = np.linalg.matrix_rank(c)
    return choi2ortho(factorize_psd_truncated(c, chi_max=rank))


def ortho2super(x):
    """transformations.py:743-749: isometry -> Choi factor -> Choi -> superoperator."""
    c = ortho2choi(x)
    return choi2super(c @ c.conj().T)


def kraus2superop(kraus_list):
    """transformations.py:170-180"""
    if not isinstance(kraus_list, list):
        kraus_list = [kraus_list]
    r, c = kraus_list[0].shape
    out = np.zeros((r * r, c * c), dtype=complex)
    for E in kraus_list:
        out += np.kron(E, E.conj())
    return out


def kraus_from_ortho(x):
    """E_k = x[k::K, :]: row a*K+k of the isometry is row a of Kraus operator k."""
    dim = x.shape[1]
    K = x.shape[0] // dim
    return [x[k::K, :] for k in range(K)]


# --------------------------------------------------------------------------------------------------
# local -> full embedding (transformations.py:405-490)
# --------------------------------------------------------------------------------------------------
def create_supertensored_from_local(localop, N, pbc=False, layer=0):
    """transformations.py:405-429"""
    assert N % 2 == 0, "Only even number of sites allowed"
    if layer % 2 == 0:
        pbc = True
    out = localop if pbc else np.eye(localop.shape[0])
    for _ in range(N // 2 - 1):
        out = np.kron(out, localop)
    return out


def get_indices_supertensored2liouvillianfull(N):
    """transformations.py:431-444"""
    src = list(chain.from_iterable((2 + 4 * i, 3 + 4 * i) for i in range((N - 2) // 2)))
    dst = list(range(N, 2 * N - 2))
    return src + [s + 2 * N for s in src], dst + [t + 2 * N for t in dst]


def swap_superop_indices(superop, src, dst, N, d, shift_pbc=False):
    """transformations.py:473-482"""
    t = np.reshape(superop, [d] * (4 * N))
    t = np.moveaxis(t, src, dst)
    if shift_pbc:
        a = list(range(N))
        b = list(range(N, 2 * N))
        idx = a[1:] + a[:1] + b[1:] + b[:1]
        t = np.transpose(t, idx + [i + 2 * N for i in idx])
    return np.reshape(t, superop.shape)


def convert_supertensored2liouvillianfull(local_tensored, N, d, shift_pbc=False):
    """transformations.py:484-490"""
    src, dst = get_indices_supertensored2liouvillianfull(N)
    return swap_superop_indices(local_tensored, src, dst, N, d, shift_pbc)


def convert_liouvillianfull2supertensored(full, N, d):
    """transformations.py:493-498"""
    dst, src = get_indices_supertensored2liouvillianfull(N)
    return swap_superop_indices(full, src, dst, N, d)


def compose_superops_list(superops):
    """transformations.py:841-851: list order = application order, M = Y_m ... Y_1."""
    out = superops[0]
    for op in superops[1:]:
        out = np.dot(op, out)
    return out


# --------------------------------------------------------------------------------------------------
# models and cost (optimization.py:56-61, 106-151)
# --------------------------------------------------------------------------------------------------
def frobenius_norm(A, B, squared=False):
    """optimization.py:56-61"""
    return np.linalg.norm(A - B, ord="fro") ** (2 if squared else 1)


def local_layers_full(xs, N, d):
    """The per-layer full-space superoperators of optimization.py:125-136 (odd list index => pbc shift)."""
    out = []
    for i, x in enumerate(xs):
        split = create_supertensored_from_local(ortho2super(x), N=N, pbc=True)
        out.append(convert_supertensored2liouvillianfull(split, N=N, d=d, shift_pbc=(i % 2 == 1)))
    return out


def model_stiefel_local(xs, N, d):
    """optimization.py:106-137"""
    return compose_superops_list(local_layers_full(xs, N, d))


def model_Ys_stiefel(xs):
    """optimization.py:147-151 (+ model_Ys :139-145, unfactorize_list :71-79)"""
    return compose_superops_list([ortho2super(x) for x in xs])


class CostSpec:
    """The closure `lambda xi: frobenius_norm(model(xi), target)` the reference notebooks build
    (e.g. experiments/trust_region_pauli_noise.ipynb cell 21), made explicit so that the closed-form
    derivative sweeps below know the computational graph."""

    def __init__(self, target, model="local", N=None, d=2):
        assert model in ("local", "full")
        self.target = np.asarray(target)
        self.model, self.N, self.d = model, N, d
        if model == "local":
            assert N is not None

    def layers(self, xs):
        if self.model == "local":
            return local_layers_full(xs, self.N, self.d)
        return [ortho2super(x) for x in xs]

    def __call__(self, xs):
        return frobenius_norm(compose_superops_list(self.layers(xs)), self.target)


# --------------------------------------------------------------------------------------------------
# restatement of the autodiff call sites (stiefel.py:423,437,450 -> egrad; stiefel.py:520 -> egrad_jvp)
# --------------------------------------------------------------------------------------------------
def _layer_vjp(spec, i, x, W):
    """Pull the cotangent W = df/dY_full(layer i) back to the isometry x (reverse of
    ortho2super [+ kron embedding + permutation])."""
    if spec.model == "local":
        N, d = spec.N, spec.d
        # undo the permutation (reverse of convert_supertensored2liouvillianfull)
        src, dst = get_indices_supertensored2liouvillianfull(N)
        t = np.reshape(W, [d] * (4 * N))
        if i % 2 == 1:
            a = list(range(N))
            b = list(range(N, 2 * N))
            idx = a[1:] + a[:1] + b[1:] + b[:1]
            perm = idx + [j + 2 * N for j in idx]
            t = np.transpose(t, np.argsort(perm))
        t = np.moveaxis(t, dst, src)
        Wt = np.reshape(t, W.shape)
        # reverse of kron(Y_loc, ..., Y_loc): product rule over the N/2 factors
        Yl = ortho2super(x)
        q = Yl.shape[0]
        nf = N // 2
        Wt = Wt.reshape([q] * nf + [q] * nf)  # (i_0..i_{nf-1}, j_0..j_{nf-1})
        Wl = np.zeros_like(Yl)
        for f in range(nf):
            cur = Wt
            # contract every factor g != f with Y_loc[i_g, j_g]
            keep_rows = list(range(nf))
            for g in reversed(range(nf)):
                if g == f:
                    continue
                r_ax = keep_rows.index(g)
                c_ax = len(keep_rows) + keep_rows.index(g)
                cur = np.tensordot(cur, Yl, axes=([r_ax, c_ax], [0, 1]))
                keep_rows.remove(g)
            Wl = Wl + cur
        Wsup = Wl
    else:
        Wsup = W
    # reverse of choi2super (an involution) and of C = c c^T, then of ortho2choi
    dim = x.shape[1]
    K = x.shape[0] // dim
    Cbar = choi2super(Wsup)
    c = ortho2choi(x)
    cbar = (Cbar + Cbar.T) @ c
    return np.reshape(cbar, [dim, dim, K]).swapaxes(1, 2).reshape(dim * K, dim)


def _layer_jvp(spec, i, x, eta):
    """Directional derivative of the full-space layer superoperator along eta (forward mode)."""
    c, cd = ortho2choi(x), ortho2choi(eta)
    Yl, Yld = choi2super(c @ c.T), choi2super(cd @ c.T + c @ cd.T)
    if spec.model == "full":
        return Yld
    N, d = spec.N, spec.d
    nf = N // 2
    tot = 0
    for f in range(nf):
        cur = Yld if f == 0 else Yl
        for g in range(1, nf):
            cur = np.kron(cur, Yld if g == f else Yl)
        tot = tot + cur
    return convert_supertensored2liouvillianfull(tot, N, d, shift_pbc=(i % 2 == 1))


def cost_and_egrad(spec, xs):
    """f and the Euclidean gradient list [df/dx_l]: what `jax.grad(func)(xi)` returns at stiefel.py:423."""
    Y = spec.layers(xs)
    m = len(Y)
    P = [Y[0]]
    for l in range(1, m):
        P.append(Y[l] @ P[-1])
    R = P[-1] - spec.target
    f = np.linalg.norm(R, "fro")
    G = np.real(R) / f  # only Re(R) carries gradient for a real model
    Z = [None] * m
    for l in range(m - 1, 0, -1):
        Z[l] = _layer_vjp(spec, l, xs[l], G @ P[l - 1].T)
        G = Y[l].T @ G
    Z[0] = _layer_vjp(spec, 0, xs[0], G)
    return f, Z


def egrad(spec, xs):
    return cost_and_egrad(spec, xs)[1]


def egrad_jvp(spec, xs, etas):
    """(egrad, d/dt egrad(x + t eta)|_0): what `jax.jvp(jax.grad(func), (x,), (eta,))` returns
    (the Euclidean part of stiefel.py:520)."""
    Y = spec.layers(xs)
    Yd = [_layer_jvp(spec, l, xs[l], etas[l]) for l in range(len(xs))]
    m = len(Y)
    P, Pd = [Y[0]], [Yd[0]]
    for l in range(1, m):
        Pd.append(Yd[l] @ P[-1] + Y[l] @ Pd[-1])
        P.append(Y[l] @ P[-1])
    R = P[-1] - spec.target
    f = np.linalg.norm(R, "fro")
    Rr = np.real(R)
    fd = np.sum(Rr * Pd[-1]) / f
    G = Rr / f
    Gd = Pd[-1] / f - Rr * (fd / f ** 2)
    Z, Zd = [None] * m, [None] * m
    for l in range(m - 1, 0, -1):
        W = G @ P[l - 1].T
        Wd = Gd @ P[l - 1].T + G @ Pd[l - 1].T
        Z[l] = _layer_vjp(spec, l, xs[l], W)
        Zd[l] = _layer_vjp_tangent(spec, l, xs[l], etas[l], W, Wd)
        Gd = Yd[l].T @ G + Y[l].T @ Gd
        G = Y[l].T @ G
    Z[0] = _layer_vjp(spec, 0, xs[0], G)
    Zd[0] = _layer_vjp_tangent(spec, 0, xs[0], etas[0], G, Gd)
    return f, Z, Zd


def _layer_vjp_tangent(spec, i, x, eta, W, Wd):
    """d/dt of _layer_vjp(spec, i, x + t eta, W + t Wd) at t = 0.  `_layer_vjp` is linear in W and
    polynomial in x, so this is Wd pulled back at x plus W pulled back along the x-perturbation; the
    second piece is taken by exact polarisation of the polynomial (degree <= N-1 in x for the local model,
    degree 1 for the full model), i.e. still closed form, no step-size error."""
    lin = _layer_vjp(spec, i, x, Wd)
    if spec.model == "full":
        # Z = (Cbar + Cbar^T) c  is linear in x for fixed W
        return lin + _layer_vjp(spec, i, eta, W)
    # local model: Z(x) = J(x)^T W with J polynomial in x.  Use complex-step-free exact differentiation:
    # differentiate the explicit formula  Wl(x) = sum_f contract_{g != f} Y_loc(x)  and  cbar = (Cb+Cb^T) c.
    N, d = spec.N, spec.d
    src, dst = get_indices_supertensored2liouvillianfull(N)
    t = np.reshape(W, [d] * (4 * N))
    if i % 2 == 1:
        a = list(range(N))
        b = list(range(N, 2 * N))
        idx = a[1:] + a[:1] + b[1:] + b[:1]
        perm = idx + [j + 2 * N for j in idx]
        t = np.transpose(t, np.argsort(perm))
    Wt = np.reshape(np.moveaxis(t, dst, src), W.shape)
    c, cd = ortho2choi(x), ortho2choi(eta)
    Yl, Yld = choi2super(c @ c.T), choi2super(cd @ c.T + c @ cd.T)
    q = Yl.shape[0]
    nf = N // 2
    Wt = Wt.reshape([q] * (2 * nf))
    Wl = np.zeros_like(Yl)
    Wld = np.zeros_like(Yl)
    for f in range(nf):
        others = [g for g in range(nf) if g != f]
        # value
        Wl += _contract_others(Wt, nf, f, {g: Yl for g in others})
        # derivative: one of the other factors replaced by Yld
        for h in others:
            Wld += _contract_others(Wt, nf, f, {g: (Yld if g == h else Yl) for g in others})
    dim = x.shape[1]
    K = x.shape[0] // dim
    Cb, Cbd = choi2super(Wl), choi2super(Wld)
    cbar_d = (Cbd + Cbd.T) @ c + (Cb + Cb.T) @ cd
    extra = np.reshape(cbar_d, [dim, dim, K]).swapaxes(1, 2).reshape(dim * K, dim)
    return lin + extra


def _contract_others(Wt, nf, f, mats):
    cur = Wt
    keep = list(range(nf))
    for g in reversed(range(nf)):
        if g == f:
            continue
        r_ax = keep.index(g)
        c_ax = len(keep) + keep.index(g)
        cur = np.tensordot(cur, mats[g], axes=([r_ax, c_ax], [0, 1]))
        keep.remove(g)
    return cur


# --------------------------------------------------------------------------------------------------
# Stiefel geometry (stiefel.py)
# --------------------------------------------------------------------------------------------------
def symmetrize(A):
    """stiefel.py:32-37"""
    return 0.5 * (A + A.T)


def antisymmetrize(A):
    """stiefel.py:40-45"""
    return 0.5 * (A - A.T)


def project(X, Z):
    """stiefel.py:23-30"""
    return Z - X @ symmetrize(X.T @ Z)


def canonical_metric(d1, d2, x):
    """stiefel.py:76-83"""
    n = x.shape[0]
    return np.trace(d1.conj().T @ (np.eye(n) - 0.5 * (x @ x.conj().T)) @ d2).real


def euclidean_metric(d1, d2):
    """stiefel.py:85-91"""
    return np.trace(d1.conj().T @ d2).real


def antisymmetric_dof(p):
    return p * (p - 1) // 2


def stiefel_tangent_dof(n, p):
    """stiefel.py:60-74"""
    return antisymmetric_dof(p) + p * (n - p)


def get_orthogonal_complement(x):
    """stiefel.py:175-190: SVD factor of I - x x^T truncated to n-p columns (basis is not unique)."""
    n, p = x.shape
    if n == p:
        return np.zeros_like(x)
    return factorize_psd_truncated(np.eye(n) - x @ x.T, chi_max=n - p, eps=1e-15)


def upper_triangular_to_antisymmetric(upper, p):
    """stiefel.py:51-58"""
    a = np.zeros((p, p))
    a[np.triu_indices(p, k=1)] = upper
    return a - a.T


def parametrization_from_tangent(x, z, x_comp=None):
    """stiefel.py:266-300 with stack=True, vectorized=True: [triu(A,1), B.ravel()]."""
    n, p = x.shape
    A = antisymmetrize(x.T @ z)  # get_antisymmetric, stiefel.py:193-202
    Au = A[np.triu_indices(p, k=1)]
    if n == p:
        return Au
    if x_comp is None:
        x_comp = get_orthogonal_complement(x)
    B = x_comp.T @ (np.eye(n) - x @ x.T) @ z  # get_arbitrary, stiefel.py:204-219
    return np.hstack([Au, B.reshape(-1)])


def parametrizations_from_vector(vec, shapes):
    """stiefel.py:221-250 (np.hsplit => equal shapes)."""
    parts = np.hsplit(vec, len(shapes))
    out = []
    for part, (n, p) in zip(parts, shapes):
        u = antisymmetric_dof(p)
        out.append((upper_triangular_to_antisymmetric(part[:u], p), part[u:].reshape(n - p, p)))
    return out


def tangent_from_parametrization(X, A, B, x_comp=None):
    """stiefel.py:302-310"""
    n, p = X.shape
    if n == p:
        return X @ A
    if x_comp is None:
        x_comp = get_orthogonal_complement(X)
    return X @ A + x_comp @ B


def polar_decomposition_rectangular(X, Z):
    """stiefel.py:323-330"""
    return scipy.linalg.polar(X + Z)[0]


def retract_stiefel(xi, eta):
    """stiefel.py:332-355 (eta is the coefficient vector)."""
    params = parametrizations_from_vector(eta, [x.shape for x in xi])
    dx = [tangent_from_parametrization(x, A, B) for x, (A, B) in zip(xi, params)]
    return [polar_decomposition_rectangular(x, z) for x, z in zip(xi, dx)]


def retract_tangent(xi, etas):
    """Same retraction with the tangent given as (n,p) matrices (basis independent form)."""
    return [polar_decomposition_rectangular(x, z) for x, z in zip(xi, etas)]


def metric_alphas(metric):
    """stiefel.py:456-461"""
    if metric == "euclidean":
        return 1.0, 1.0
    if metric == "canonical":
        return 1.0, 0.5
    if isinstance(metric, (tuple, list)) and len(metric) == 2:   # general (alpha0, alpha1), gradient_stiefel_general :411-426
        return float(metric[0]), float(metric[1])
    raise ValueError(f"{metric} is not a valid metric value")


def gradient_ambient2riemannian(x, g, alpha0=1, alpha1=1):
    """stiefel.py:398-409"""
    return (1 / alpha0) * g + 0.5 * ((1 / alpha1) - (2 / alpha0)) * x @ x.T @ g - 0.5 * (1 / alpha1) * x @ g.T @ x


def riemannian_connection(D_nu, nu, eta, x, alpha0=1, alpha1=1):
    """stiefel.py:468-475"""
    In = np.eye(x.shape[0])
    return D_nu + 0.5 * x @ (eta.T @ nu + nu.T @ eta) + ((alpha0 - alpha1) / alpha0) * (In - x @ x.T) @ (
        eta @ nu.T + nu @ eta.T) @ x


def rgrad(spec, xs, metric="euclidean"):
    """gradient_stiefel_general, stiefel.py:411-426 -> list of (n,p) tangent matrices."""
    a0, a1 = metric_alphas(metric)
    return [gradient_ambient2riemannian(x, z, a0, a1) for x, z in zip(xs, egrad(spec, xs))]


def cost_rgrad(spec, xs, metric="euclidean"):
    a0, a1 = metric_alphas(metric)
    f, Z = cost_and_egrad(spec, xs)
    return f, [gradient_ambient2riemannian(x, z, a0, a1) for x, z in zip(xs, Z)]


def gradient_stiefel_vec(xs, spec, metric="euclidean"):
    """stiefel.py:453-466 -> coefficient vector."""
    return np.hstack([parametrization_from_tangent(x, g) for x, g in zip(xs, rgrad(spec, xs, metric))])


def _rgrad_jvp(x, eta, Z, Zd, a0, a1):
    """Directional derivative of the map x -> gradient_ambient2riemannian(x, Z(x)) (product rule on
    stiefel.py:409)."""
    c1 = 0.5 * ((1 / a1) - (2 / a0))
    c2 = 0.5 * (1 / a1)
    return (Zd / a0 + c1 * (eta @ (x.T @ Z) + x @ (eta.T @ Z) + x @ (x.T @ Zd))
            - c2 * (eta @ (Z.T @ x) + x @ (Zd.T @ x) + x @ (Z.T @ eta)))


def triple(spec, xs, etas, metric="canonical"):
    """Unit of work U1: (f, rgrad, H eta) sharing the primal pass.  H eta is one column-combination of
    the matrix stiefel.py:498-529 builds, returned as (n,p) tangent matrices:
        jvp of grad_func (:520) -> riemannian_connection (:524) -> tangent projection (what
        parametrization_from_tangent :527 keeps)."""
    a0, a1 = metric_alphas(metric)
    f, Z, Zd = egrad_jvp(spec, xs, etas)
    g, h = [], []
    for x, eta, z, zd in zip(xs, etas, Z, Zd):
        nu = gradient_ambient2riemannian(x, z, a0, a1)
        dnu = _rgrad_jvp(x, eta, z, zd, a0, a1)
        conn = riemannian_connection(dnu, nu, eta, x, a0, a1)
        g.append(nu)
        h.append(project(x, conn))
    return f, g, h


def hvp(spec, xs, etas, metric="canonical"):
    return triple(spec, xs, etas, metric)[2]


class HessianOperator:
    """Matrix-free stand-in for the ndarray `riemannian_hessian_vec` returns (stiefel.py:534-544); the TR loop
    only ever applies `@` to a 1-D coefficient vector (trust_region_rcopt.py:67,116)."""

    def __init__(self, spec, xs, metric="canonical"):
        self.spec, self.xs, self.metric = spec, [np.asarray(x) for x in xs], metric
        self.comps = [get_orthogonal_complement(x) for x in self.xs]
        self.shapes = [x.shape for x in self.xs]
        self.n_matvec = 0

    def __matmul__(self, v):
        self.n_matvec += 1
        params = parametrizations_from_vector(np.asarray(v), self.shapes)
        etas = [tangent_from_parametrization(x, A, B, xc) for x, (A, B), xc in zip(self.xs, params, self.comps)]
        h = hvp(self.spec, self.xs, etas, self.metric)
        return np.hstack([parametrization_from_tangent(x, z, xc) for x, z, xc in zip(self.xs, h, self.comps)])

    def to_dense(self):
        size = sum(stiefel_tangent_dof(*s) for s in self.shapes)
        eye = np.eye(size)
        return np.stack([self @ eye[:, k] for k in range(size)], axis=1)


def riemannian_hessian_vec(xs, spec, metric="euclidean"):
    """Dense matrix, built one column per elementary tangent direction exactly like stiefel.py:498-544."""
    return HessianOperator(spec, xs, metric).to_dense()


# --------------------------------------------------------------------------------------------------
# trust region (trust_region_rcopt.py)
# --------------------------------------------------------------------------------------------------
def solve_quadratic_equation(p, q):
    """trust_region_rcopt.py:155-166"""
    if p ** 2 - q < 0:
        raise ValueError("require non-negative discriminant")
    if p == 0:
        x = np.sqrt(-q)
        return (-x, x)
    x1 = -(p + np.sign(p) * np.sqrt(p ** 2 - q))
    x2 = q / x1
    return tuple(sorted((x1, x2)))


def _move_to_boundary(b, d, radius):
    """trust_region_rcopt.py:137-152"""
    dsq = np.dot(d, d)
    if dsq == 0:
        warnings.warn("input vector 'd' is zero")
        return b
    p = np.dot(b, d) / dsq
    q = (np.dot(b, b) - radius ** 2) / dsq
    t = solve_quadratic_equation(p, q)[1]
    if t < 0:
        warnings.warn("encountered t < 0")
    return t


def truncated_cg(grad, hess, radius, **kwargs):
    """trust_region_rcopt.py:92-134"""
    maxiter = kwargs.get("maxiter", 2 * len(grad))
    abstol = kwargs.get("abstol", 1e-8)
    reltol = kwargs.get("reltol", 1e-6)
    r = grad.copy()
    rsq = np.dot(r, r)
    stoptol = max(abstol, reltol * np.sqrt(rsq))
    z = np.zeros_like(r)
    d = -r
    for _ in range(maxiter):
        Hd = hess @ d
        dHd = np.dot(d, Hd)
        t = _move_to_boundary(z, d, radius)
        alpha = rsq / dHd
        if dHd <= 0 or alpha > t:
            return z + t * d, True
        r += alpha * Hd
        z += alpha * d
        rsq_next = np.dot(r, r)
        if np.sqrt(rsq_next) <= stoptol:
            return z, False
        beta = rsq_next / rsq
        d = -r + beta * d
        rsq = rsq_next
    return z, False


def riemannian_trust_region_optimize(f, retract, gradfunc, hessfunc, x_init, save_x=False, **kwargs):
    """trust_region_rcopt.py:5-89 (prints and the KeyboardInterrupt handler dropped; also returns a
    per-iteration log used by the per-step parity tests)."""
    rho_trust = kwargs.get("rho_trust", 0.125)
    radius = kwargs.get("radius_init", 0.01)
    maxradius = kwargs.get("maxradius", 0.1)
    niter = kwargs.get("niter", 20)
    tcg_kwargs = {k: kwargs["tcg_" + k] for k in ("maxiter", "abstol", "reltol") if "tcg_" + k in kwargs}
    assert 0 <= rho_trust < 0.25
    x = x_init
    f_iter, x_iter, log_ = [], [x], []
    for k in range(niter):
        grad = gradfunc(x)
        hess = hessfunc(x)
        eta, on_boundary = truncated_cg(grad, hess, radius, **tcg_kwargs)
        x_next = retract(x, eta)
        fx = f(x)
        f_iter.append(fx)
        rho = (f(x_next) - fx) / (np.dot(grad, eta) + 0.5 * np.dot(eta, hess @ eta))
        if rho < 0.25:
            radius *= 0.25
        elif rho > 0.75 and on_boundary:
            radius = min(2 * radius, maxradius)
        accepted = rho > rho_trust
        if accepted:
            x = x_next
        log_.append(dict(rho=rho, radius=radius, on_boundary=on_boundary, accepted=accepted,
                         n_matvec=getattr(hess, "n_matvec", None)))
        if save_x or k == niter - 1:
            x_iter.append(x)
    return x_iter, f_iter, radius, log_


def trust_region_fit(spec, xs0, metric="canonical", **kwargs):
    """The notebook driver (experiments/trust_region_pauli_noise.ipynb cells 21, 32-33) with the matrix-free
    Hessian operator in place of the dense matrix."""
    return riemannian_trust_region_optimize(
        spec, retract_stiefel, lambda xi: gradient_stiefel_vec(xi, spec, metric),
        lambda xi: HessianOperator(spec, xi, metric), list(xs0), **kwargs)


# --------------------------------------------------------------------------------------------------
# tangent-matrix (basis independent) trust region: what the batched device solver computes
# --------------------------------------------------------------------------------------------------
def canonical_dot(xs, a, b):
    """sum_l tr(a_l^T (I - x_l x_l^T / 2) b_l) == Euclidean dot of the (A-upper, B) coefficient vectors."""
    tot = 0.0
    for x, u, v in zip(xs, a, b):
        tot += np.sum(u * v) - 0.5 * np.sum((x.T @ u) * (x.T @ v))
    return tot


# --------------------------------------------------------------------------------------------------
# synthetic workloads
# --------------------------------------------------------------------------------------------------
def random_stiefel(n, p, rng):
    """x = qf(G), G ~ N(0,1)^{n x p}, sign-fixed QR (diag R > 0)."""
    q, r = np.linalg.qr(rng.standard_normal((n, p)))
    return q * np.sign(np.diag(r))


def kitaev_fullsites_target(N=4, d=2, gamma=1.0, tau=4.0, pbc=False):
    """Target of experiments/trust_region_fullsites_ansatz.ipynb cell 1."""
    Lvec, _, _ = create_2local_liouvillians(get_kitaev_nn_linbladian(gamma), N, d, pbc)
    return exp_operator_dt(Lvec, tau)


def c3_instance(b, m=3, n=256, p=16):
    """Instance b of BASELINE config 3: random initial point and unit-canonical-norm random tangent."""
    rng = np.random.default_rng(1234 + b)
    xs = [random_stiefel(n, p, rng) for _ in range(m)]
    rng2 = np.random.default_rng(10 ** 6 + b)
    etas = [project(x, rng2.standard_normal((n, p))) for x in xs]
    nrm = np.sqrt(canonical_dot(xs, etas, etas))
    return xs, [e / nrm for e in etas]


# --------------------------------------------------------------------------------------------------
# counter-based random numbers of the device factory (csrc/extras.inl) restated in NumPy
# --------------------------------------------------------------------------------------------------
def philox4x32_10(c, k):
    """Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; Random123 constants) on arrays of
    counters c (..., 4) and keys k (..., 2), dtype uint32 -> (..., 4) uint32."""
    c = [np.asarray(c[..., i], dtype=np.uint64) for i in range(4)]
    k0, k1 = np.asarray(k[..., 0], dtype=np.uint64), np.asarray(k[..., 1], dtype=np.uint64)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c[0]
        p1 = np.uint64(0xCD9E8D57) * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ k0, p1 & mask, (p0 >> np.uint64(32)) ^ c[3] ^ k1, p0 & mask]
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return np.stack(c, axis=-1).astype(np.uint32)


def philox_normal(seed, count, stream=0):
    """Standard normals e = 0..count-1 of one instance under the device convention: counter (e_lo, e_hi, stream, 0), key
    (seed_lo, seed_hi); u = ((x_hi << 32 | x_lo) >> 11 + 0.5) 2^-53 from each half of the output; Box-Muller cosine branch."""
    e = np.arange(count, dtype=np.uint64)
    ctr = np.stack([e & np.uint64(0xFFFFFFFF), e >> np.uint64(32), np.full(count, stream, dtype=np.uint64),
                    np.zeros(count, dtype=np.uint64)], axis=-1)
    seed = np.uint64(seed)
    key = np.broadcast_to(np.array([seed & np.uint64(0xFFFFFFFF), seed >> np.uint64(32)], dtype=np.uint64), (count, 2))
    x = philox4x32_10(ctr, key).astype(np.uint64)
    u1 = (((x[:, 0] << np.uint64(32)) | x[:, 1]) >> np.uint64(11)).astype(np.float64) + 0.5
    u2 = (((x[:, 2] << np.uint64(32)) | x[:, 3]) >> np.uint64(11)).astype(np.float64) + 0.5
    u1 *= 2.0 ** -53
    u2 *= 2.0 ** -53
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def kicked_start(xs0, seed, kick=1e-2, stream=2):
    """Trotter start + random tangent kick of canonical norm `kick` (what otn_kicked_starts computes on the device)."""
    xs0 = [np.asarray(x, dtype=np.float64) for x in xs0]
    n, p = xs0[0].shape
    g = philox_normal(seed, len(xs0) * n * p, stream).reshape(len(xs0), n, p)
    etas = [project(x, gi) for x, gi in zip(xs0, g)]
    nrm = np.sqrt(canonical_dot(xs0, etas, etas))
    return retract_tangent(xs0, [kick * e / nrm for e in etas])
