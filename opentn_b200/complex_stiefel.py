"""Complex Stiefel manifold: the extension `opentn.stiefel` marks "TODO: add back the conj" (stiefel.py:29,36,44,184,201,214,218,320),
over the complex handle of libotn (OTN_FLAG_COMPLEX, csrc/complex.cuh).

`closures(fit, metric)` returns the four callables `riemannian_trust_region_optimize` expects (trust_region_rcopt.py:5) for a
list of COMPLEX isometries.  Cost, Riemannian gradient, Hessian-vector products and the polar retraction run on the GPU; the
coefficient vectors the optimiser works with are host glue, like in the real mirror (`opentn_b200.stiefel`):

    tangent  z = x A + x_perp B,   A skew-Hermitian (p^2 real degrees of freedom),  B in C^{(n-p) x p}
    vector   [Re A_ij, Im A_ij (i<j), Im A_ii / sqrt2, Re B, Im B]       (stiefel.py:266-300 with the conjugates put back)

so that the Euclidean dot of two vectors is the canonical metric Re tr(d1^H (I - x x^H / 2) d2) -- the inner product the
reference's truncated CG uses implicitly (trust_region_rcopt.py:111-131).
"""
from __future__ import annotations

import numpy as np

from . import _lib
from ._lib import check, ptr
from .fit import StiefelFit, metric_alphas
from . import transformations as _T


def get_orthogonal_complement(x):
    """Orthonormal basis of range(x)^perp for a complex isometry (complete QR), stiefel.py:175-190."""
    x = np.asarray(x, dtype=np.complex128)
    n, p = x.shape
    q, _ = np.linalg.qr(x, mode="complete")
    return np.ascontiguousarray(q[:, p:])


def parametrization_from_tangent(x, z, x_comp=None):
    x, z = np.asarray(x, dtype=np.complex128), np.asarray(z, dtype=np.complex128)
    n, p = x.shape
    A = x.conj().T @ z
    A = 0.5 * (A - A.conj().T)
    iu = np.triu_indices(p, k=1)
    parts = [A[iu].real, A[iu].imag, np.diag(A).imag / np.sqrt(2.0)]
    if n > p:
        xc = get_orthogonal_complement(x) if x_comp is None else x_comp
        B = xc.conj().T @ z
        parts += [B.real.ravel(), B.imag.ravel()]
    return np.concatenate(parts)


def tangent_from_parametrization(x, vec, x_comp=None):
    x = np.asarray(x, dtype=np.complex128)
    n, p = x.shape
    nu = p * (p - 1) // 2
    iu = np.triu_indices(p, k=1)
    A = np.zeros((p, p), dtype=np.complex128)
    A[iu] = vec[:nu] + 1j * vec[nu:2 * nu]
    A = A - A.conj().T
    A[np.diag_indices(p)] = 1j * np.sqrt(2.0) * vec[2 * nu:2 * nu + p]
    z = x @ A
    if n > p:
        xc = get_orthogonal_complement(x) if x_comp is None else x_comp
        rest, half = vec[2 * nu + p:], (n - p) * p
        z = z + xc @ (rest[:half] + 1j * rest[half:]).reshape(n - p, p)
    return z


def tangent_dof(n, p):
    return p * p + 2 * (n - p) * p


def polar_decomposition_rectangular(X, Z):
    """Polar factor of X + Z for complex (stacks of) matrices on the GPU (otn_retract_complex), stiefel.py:323-330."""
    xb = np.ascontiguousarray(X, dtype=np.complex128)
    lead, (n, p) = xb.shape[:-2], xb.shape[-2:]
    xb = xb.reshape(-1, n, p)
    zb = np.ascontiguousarray(Z, dtype=np.complex128).reshape(xb.shape)
    out = np.empty_like(xb)
    status = np.zeros(xb.shape[0], dtype=np.int32)
    check(_lib.load().otn_retract_complex(ptr(xb), ptr(zb), xb.shape[0], n, p, ptr(out), ptr(status), _T._DEVICE))
    if status.any():
        raise np.linalg.LinAlgError("polar retraction: Newton-Schulz did not converge ((X+Z)^H (X+Z) too far from the identity)")
    return out.reshape(lead + (n, p))


class _Basis:
    """Complement bases of one iterate, shared by gradient, Hessian and retraction."""

    def __init__(self, xs):
        self.xs = [np.asarray(x, dtype=np.complex128) for x in xs]
        self.comps = [get_orthogonal_complement(x) if x.shape[0] > x.shape[1] else None for x in self.xs]
        self.sizes = [tangent_dof(*x.shape) for x in self.xs]

    def to_vector(self, zs):
        return np.concatenate([parametrization_from_tangent(x, z, c) for x, z, c in zip(self.xs, zs, self.comps)])

    def to_tangents(self, vec):
        out, o = [], 0
        for x, c, sz in zip(self.xs, self.comps, self.sizes):
            out.append(tangent_from_parametrization(x, vec[o:o + sz], c))
            o += sz
        return out


class ComplexHessianOperator:
    """`hess @ v` for the host trust-region loop: one cached-primal Hessian-vector product on the GPU per application."""

    def __init__(self, fit: StiefelFit, xs, basis: _Basis):
        self.fit, self.basis = fit, basis
        self.xs = np.stack(basis.xs)
        self.size = sum(basis.sizes)
        self.shape = (self.size, self.size)
        self.n_matvec = 0
        fit.cost_grad(self.xs)
        self._token = fit._cached[2]

    def __matmul__(self, v):
        if self.fit._cached is None or self.fit._cached[2] is not self._token:
            self.fit.cost_grad(self.xs)
            self._token = self.fit._cached[2]
        etas = np.stack(self.basis.to_tangents(np.asarray(v, dtype=np.float64)))
        self.n_matvec += 1
        return self.basis.to_vector(list(self.fit.hvp_cached(etas)))


def closures(fit: StiefelFit, metric=None):
    """(f, retract, gradfunc, hessfunc) for `riemannian_trust_region_optimize` on complex isometries (model='full')."""
    if metric is not None and metric_alphas(metric) != metric_alphas(fit.metric):
        fit.set_metric(metric)
    cache = {}

    def basis_of(xs):
        key = tuple(np.asarray(x).tobytes() for x in xs)
        if cache.get("key") != key:
            cache["key"], cache["basis"] = key, _Basis(xs)
        return cache["basis"]

    def f(xs):
        return fit.cost(np.stack([np.asarray(x, dtype=np.complex128) for x in xs]))

    def gradfunc(xs):
        b = basis_of(xs)
        _, g = fit.cost_grad(np.stack(b.xs))
        return b.to_vector(list(g))

    def hessfunc(xs):
        return ComplexHessianOperator(fit, xs, basis_of(xs))

    def retract(xs, eta):
        b = basis_of(xs)
        return list(polar_decomposition_rectangular(np.stack(b.xs), np.stack(b.to_tangents(np.asarray(eta, dtype=np.float64)))))

    return f, retract, gradfunc, hessfunc
