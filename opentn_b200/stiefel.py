"""Host-side mirror of `opentn.stiefel` (same names and argument meaning) over libotn.

The reference differentiates an opaque Python closure with JAX (`stiefel.py:423,437,450,520`).  Here `func` must be an
`opentn_b200.StiefelFit`; anything else raises `TypeError` (there is no CPU fallback).

Coefficient vectors.  `gradient_stiefel_vec`, the Hessian operator and `retract_stiefel` speak the reference's
(A-upper, B) coefficient basis (`stiefel.py:266-300`), which needs an orthonormal complement X_perp.  That basis is not
unique (the reference takes it from an SVD, `stiefel.py:190`); any orthonormal X_perp gives the same tangent matrices,
the same coefficient-space dot products and therefore the same trust-region iterates.  X_perp is host glue (one QR per
isometry, cached per x); all arithmetic on (n,p) matrices runs on the GPU.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from ._lib import as_f64, check, ptr
from .fit import StiefelFit, metric_alphas, metric_id
from . import transformations as _T


def _dev():
    return _T._DEVICE


# ---- metrics / checks (diagnostics, host) -------------------------------------------------------------------------------
def symmetrize(A):
    """stiefel.py:32-37"""
    return 0.5 * (A + A.T)


def antisymmetrize(A):
    """stiefel.py:40-45"""
    return 0.5 * (A - A.T)


def antisymmetric_dof(x=None, p=None) -> int:
    """stiefel.py:60-64"""
    if not p:
        p = x.shape[1]
    return p * (p - 1) // 2


def arbitrary_dof(x=None, n=None, p=None) -> int:
    """stiefel.py:66-70"""
    if not n or not p:
        n, p = x.shape
    return p * (n - p)


def stiefel_tangent_dof(x) -> int:
    """stiefel.py:72-74"""
    return antisymmetric_dof(x) + arbitrary_dof(x)


def tuple2int(i, j, cols):
    """stiefel.py:93-96"""
    assert 0 <= j < cols, "j out of range"
    return i * cols + j


def int2tuple(k, cols):
    """stiefel.py:98-100"""
    return k // cols, k % cols


def get_k_unit_matrix(dim0, dim1, k=0):
    """stiefel.py:102-107"""
    assert 0 <= k < dim0 * dim1, f"k:{k} value is out of range. Should be in f[0,{dim0 * dim1})"
    E = np.zeros(dim0 * dim1)
    E[k] = 1.0
    return E.reshape(dim0, dim1)


def get_ij_unit_matrix(dim0, dim1, i=0, j=0):
    """stiefel.py:109-111"""
    return get_k_unit_matrix(dim0, dim1, tuple2int(i, j, dim1))


def euclidean_metric(delta1, delta2) -> float:
    """stiefel.py:85-91"""
    return float(np.sum(np.asarray(delta1).conj() * np.asarray(delta2)).real)


def canonical_metric(delta1, delta2, x) -> float:
    """tr(d1^T (I - x x^T / 2) d2), stiefel.py:76-83 (GPU)."""
    x = as_f64(x)
    n, p = x.shape
    out = np.empty(1)
    check(_lib.load().otn_canonical_dot(ptr(x), ptr(as_f64(delta1)), ptr(as_f64(delta2)), 1, 1, n, p, ptr(out), _dev()))
    return float(out[0])


def is_isometry_2(x) -> bool:
    """stiefel.py:357-359"""
    return bool(np.allclose(x.conj().T @ x, np.eye(x.shape[1])))


def check_isometries(x_list, show_idx=False):
    """stiefel.py:361-372"""
    ok = [is_isometry_2(op) for op in x_list]
    if not show_idx:
        return ok
    bad = [i for i, v in enumerate(ok) if not v]
    if not bad:
        print("all elements are isometries")
        return None
    print("indices which are not isometries")
    return bad


is_isometry = check_isometries


def is_in_tangent_space(X, Z) -> bool:
    """stiefel.py:384-388"""
    assert X.ndim == Z.ndim == 2, "only matrices allowed"
    return bool(np.allclose(X.conj().T @ Z, -Z.conj().T @ X))


# ---- GPU manifold primitives ----------------------------------------------------------------------------------------
def _batched(x):
    x = as_f64(x)
    lead = x.shape[:-2]
    n, p = x.shape[-2:]
    return x.reshape(-1, n, p), lead, n, p


def project(X, Z):
    """Z - X sym(X^T Z), stiefel.py:23-30 (accepts stacks (..., n, p))."""
    xb, lead, n, p = _batched(X)
    zb = as_f64(Z).reshape(xb.shape)
    out = np.empty_like(xb)
    check(_lib.load().otn_project(ptr(xb), ptr(zb), xb.shape[0], n, p, ptr(out), _dev()))
    return out.reshape(lead + (n, p))


def gradient_ambient2riemannian(x, gradient, alpha0=1, alpha1=1):
    """stiefel.py:398-409"""
    xb, lead, n, p = _batched(x)
    gb = as_f64(gradient).reshape(xb.shape)
    out = np.empty_like(xb)
    check(_lib.load().otn_rgrad_map(ptr(xb), ptr(gb), xb.shape[0], n, p, float(alpha0), float(alpha1), ptr(out), _dev()))
    return out.reshape(lead + (n, p))


def polar_decomposition_rectangular(X, Z):
    """Polar factor of X + Z, stiefel.py:323-330 (accepts stacks)."""
    xb, lead, n, p = _batched(X)
    zb = as_f64(Z).reshape(xb.shape)
    out = np.empty_like(xb)
    status = np.zeros(xb.shape[0], dtype=np.int32)
    check(_lib.load().otn_retract(ptr(xb), ptr(zb), xb.shape[0], n, p, ptr(out), ptr(status), _dev()))
    if status.any():
        raise np.linalg.LinAlgError("polar retraction: (X+Z)^T (X+Z) is not positive definite")
    return out.reshape(lead + (n, p))


def riemannian_connection(D_nu, nu, eta, x, alpha0=1, alpha1=1):
    """stiefel.py:468-475 (GPU; accepts stacks).  Inside the accelerated HVP the connection is fused into the epilogue
    kernel (csrc/stiefel.cuh: riem_mma_kernel)."""
    xb, lead, n, p = _batched(x)
    out = np.empty_like(xb)
    check(_lib.load().otn_connection(ptr(as_f64(D_nu).reshape(xb.shape)), ptr(as_f64(nu).reshape(xb.shape)),
                                     ptr(as_f64(eta).reshape(xb.shape)), ptr(xb), xb.shape[0], n, p, float(alpha0), float(alpha1),
                                     ptr(out), _dev()))
    return out.reshape(lead + (n, p))


# ---- coefficient basis (host glue) --------------------------------------------------------------------------------------
_COMP_CACHE = {}


def get_orthogonal_complement(x):
    """An orthonormal basis of range(x)^perp, stiefel.py:175-190 (zeros for the square case, like the reference).
    The basis is not unique; a complete QR is used and cached by the bytes of x so that gradient, Hessian and
    retraction of one iterate share it."""
    x = np.asarray(x, dtype=np.float64)
    n, p = x.shape
    if n == p:
        return np.zeros_like(x)
    key = (x.shape, x.tobytes())
    xc = _COMP_CACHE.get(key)
    if xc is None:
        q, _ = np.linalg.qr(x, mode="complete")
        xc = np.ascontiguousarray(q[:, p:])
        if len(_COMP_CACHE) > 256:
            _COMP_CACHE.clear()
        _COMP_CACHE[key] = xc
    return xc


def square_to_upper_triangular(a):
    """stiefel.py:47-49"""
    return a[np.triu_indices_from(a, k=1)]


def upper_triangular_to_antisymmetric(upper, p):
    """stiefel.py:51-58"""
    assert upper.ndim == 1
    assert len(upper) == antisymmetric_dof(p=p), f"length of upper:{len(upper)} does not match with p:{p}"
    a = np.zeros((p, p))
    a[np.triu_indices(p, k=1)] = upper
    return a - a.T


def get_antisymmetric(X, Z):
    """stiefel.py:193-202"""
    return antisymmetrize(X.T @ Z)


def get_arbitrary(X_comp, Z, X=None, Px_comp=None):
    """stiefel.py:204-219"""
    if Px_comp is None and X is None:
        raise ValueError("at least one of X or Px_comp should be given")
    if Px_comp is None:
        return X_comp.T @ (Z - X @ (X.T @ Z))
    return X_comp.T @ Px_comp @ Z


def parametrization_from_tangent(x, z, x_comp=None, stack=True, vectorized=True):
    """(A, B) with z = x A + x_perp B, stiefel.py:266-300."""
    x, z = np.asarray(x), np.asarray(z)
    n, p = x.shape
    A = get_antisymmetric(x, z)
    if n == p:
        return square_to_upper_triangular(A) if vectorized else A
    if x_comp is None:
        x_comp = get_orthogonal_complement(x)
    B = get_arbitrary(x_comp, z, X=x)
    if not stack:
        return A, B
    if vectorized:
        return np.hstack([square_to_upper_triangular(A), B.reshape(-1)])
    return np.vstack([A, B])


def parametrizations_from_vector(param_vec, shapes):
    """stiefel.py:221-250 (np.hsplit => all isometries share a shape)."""
    parts = np.hsplit(np.asarray(param_vec), len(shapes))
    out = []
    for part, (n, p) in zip(parts, shapes):
        u = antisymmetric_dof(p=p)
        out.append((upper_triangular_to_antisymmetric(part[:u], p=p), _T.unvectorize(part[u:], n - p, p) if n != p
                    else np.zeros((0, p))))
    return out


def tangent_from_parametrization(X, A, B):
    """stiefel.py:302-310"""
    n, p = X.shape
    assert A.shape == (p, p), "Wrong shape for A"
    if n == p:
        return X @ A
    assert B.shape == (n - p, p), "Wrong shape for B"
    return X @ A + get_orthogonal_complement(X) @ B


def parametrization_from_tangent_square(x, z, vectorized=True):
    """A with z = x A on the orthogonal group (n == p), stiefel.py:252-264."""
    A = get_antisymmetric(np.asarray(x), np.asarray(z))
    return square_to_upper_triangular(A) if vectorized else A


# ---- elementary tangent directions (the columns of the reference's dense Hessian are indexed by them) ------------------
def get_unit_matrices(ops):
    """One E_00 unit matrix per operator shape, stiefel.py:112-121."""
    return [get_ij_unit_matrix(*np.shape(op)) for op in ops]


def get_elementary_antisymmetric(k, x):
    """x (E_ij - E_ji), (i < j) the k-th upper-triangle position, stiefel.py:123-136."""
    x = np.asarray(x)
    upper = np.zeros(antisymmetric_dof(x))
    upper[k] = 1.0
    return x @ upper_triangular_to_antisymmetric(upper, p=x.shape[1])


def get_elementary_arbitrary(k, x):
    """x_perp E_k with E_k the k-th (n-p) x p unit matrix, stiefel.py:139-148."""
    x = np.asarray(x)
    n, p = x.shape
    return get_orthogonal_complement(x) @ get_k_unit_matrix(n - p, p, k)


def get_elementary_tangent_direction(k, x):
    """k-th basis direction of T_x St(n,p): antisymmetric block first, then the complement block, stiefel.py:150-173."""
    x = np.asarray(x)
    assert 0 <= k < stiefel_tangent_dof(x), "k it is out of range"
    na = antisymmetric_dof(x)
    return get_elementary_antisymmetric(k, x) if k < na else get_elementary_arbitrary(k - na, x)


# ---- predicates and small utilities ---------------------------------------------------------------------------------------
def is_hermitian(H):
    """True, or (False, ||H - H^+||) when it is not (atol 1e-8), stiefel.py:376-382."""
    H = np.asarray(H)
    if np.allclose(H.conj().T, H, atol=1e-8):
        return True
    return False, np.linalg.norm(H - H.conj().T)


def is_antisymmetric(A):
    """stiefel.py:390-392"""
    A = np.asarray(A)
    return bool(np.allclose(A, -A.conj().T))


def is_symmetric(C):
    """stiefel.py:394-396"""
    C = np.asarray(C)
    return bool(np.allclose(C, C.conj().T))


def random_psd(dim, normalized=True):
    """Random PSD matrix B B^T from uniform entries (trace 1 if normalised), stiefel.py:15-21."""
    b = np.random.rand(dim, dim)
    psd = b @ b.conj().T
    return psd / np.trace(psd) if normalized else psd


def polar_decomposition_stiefel(X, Z):
    """(X + Z)(1 + Z^T Z)^(-1/2) for a TANGENT Z (Absil eq. 4.7), stiefel.py:312-321.  For a tangent vector X^T Z is skew, so
    (X+Z)^T (X+Z) = 1 + Z^T Z and this is the polar factor `polar_decomposition_rectangular` computes on the device."""
    X, Z = np.asarray(X), np.asarray(Z)
    assert X.shape == Z.shape, "shapes don't match"
    return polar_decomposition_rectangular(X, Z)


def _tangents_from_vector(xi, vec):
    params = parametrizations_from_vector(vec, [np.asarray(x).shape for x in xi])
    return [tangent_from_parametrization(np.asarray(x), A, B) for x, (A, B) in zip(xi, params)]


def _vector_from_tangents(xi, zs):
    return np.hstack([parametrization_from_tangent(np.asarray(x), np.asarray(z)) for x, z in zip(xi, zs)])


def retract_stiefel(xi, eta):
    """Polar retraction of the coefficient vector eta at the list of isometries xi, stiefel.py:332-355."""
    dx = np.stack(_tangents_from_vector(xi, eta))
    return list(polar_decomposition_rectangular(np.stack([np.asarray(x) for x in xi]), dx))


# ---- gradients / Hessian of a StiefelFit ---------------------------------------------------------------------------
def _need_fit(func) -> StiefelFit:
    if not isinstance(func, StiefelFit):
        raise TypeError("the accelerated path differentiates opentn_b200.StiefelFit cost objects only; arbitrary Python "
                        "closures are not supported (no CPU fallback)")
    return func


def gradient_stiefel_general(xi, func, alpha0=1, alpha1=1):
    """Riemannian gradient list for ANY metric (alpha0, alpha1) > 0 of the family, stiefel.py:411-426."""
    fit = _need_fit(func)
    metric = (float(alpha0), float(alpha1))
    old = fit.metric
    same = metric_alphas(old) == metric
    if not same:
        fit.set_metric(metric)
    try:
        _, g = fit.cost_grad(xi)
    finally:
        if not same:
            fit.set_metric(old)
    return list(g)


def gradient_stiefel(xi, func):
    """stiefel.py:428-439"""
    return gradient_stiefel_general(xi, func, 1, 1)


def gradient_canonical(xi, func):
    """stiefel.py:441-451"""
    return gradient_stiefel_general(xi, func, 1, 0.5)


def gradient_stiefel_vec(xi, func, metric="euclidean"):
    """Coefficient vector of the Riemannian gradient, stiefel.py:453-466."""
    a0, a1 = metric_alphas(metric)
    return _vector_from_tangents(xi, gradient_stiefel_general(xi, func, a0, a1))


class RiemannianHessianOperator:
    """What `riemannian_hessian_vec` returns here: a matrix-free operator.  The reference's TR loop only ever applies
    `hess @ vector` (trust_region_rcopt.py:67,116); `.to_dense()` builds the reference's dense matrix column by column
    (stiefel.py:498-544) for tests."""

    def __init__(self, x, fit: StiefelFit, metric: str):
        self.x = [np.asarray(a, dtype=np.float64) for a in x]
        self.fit, self.metric = fit, metric
        self.size = sum(stiefel_tangent_dof(a) for a in self.x)
        self.shape = (self.size, self.size)
        self.n_matvec = 0
        self._token = None
        self._ensure_cache()

    def _ensure_cache(self):
        """(Re)build the device-side primal cache if another evaluation used the handle since ours."""
        fit = self.fit
        cached = fit._cached
        if cached is not None and cached[2] is self._token and metric_alphas(fit.metric) == metric_alphas(self.metric):
            return
        if metric_alphas(fit.metric) != metric_alphas(self.metric):
            fit.set_metric(self.metric)
        fit.cost_grad(self.x)
        self._token = fit._cached[2]

    def __matmul__(self, v):
        v = np.asarray(v, dtype=np.float64)
        if v.ndim != 1 or v.shape[0] != self.size:
            raise ValueError(f"expected a vector of length {self.size}")
        self._ensure_cache()
        etas = np.stack(_tangents_from_vector(self.x, v))
        h = self.fit.hvp_cached(etas)
        self.n_matvec += 1
        return _vector_from_tangents(self.x, list(h))

    def to_dense(self, chunk: int = 512):
        """The reference's dense Hessian (stiefel.py:498-529 builds it with one un-jitted `jax.jvp(grad)` per tangent degree of
        freedom: 14.6 s for the 450 columns of config 1).  Here the columns are a batch: `chunk` elementary directions at a time
        go through one batched cost+grad+HVP evaluation at the replicated point."""
        xs = np.stack(self.x)
        out = np.empty((self.size, self.size))
        if metric_alphas(self.fit.metric) != metric_alphas(self.metric):
            self.fit.set_metric(self.metric)
        e = np.zeros(self.size)
        for c0 in range(0, self.size, chunk):
            c1 = min(self.size, c0 + chunk)
            etas = np.empty((c1 - c0,) + xs.shape)
            for k in range(c0, c1):
                e[k] = 1.0
                etas[k - c0] = np.stack(_tangents_from_vector(self.x, e))
                e[k] = 0.0
            _, _, h = self.fit.triple(np.broadcast_to(xs, etas.shape).copy(), etas)
            for k in range(c0, c1):
                out[:, k] = _vector_from_tangents(self.x, list(h[k - c0]))
        self.n_matvec += self.size
        return out

    def __array__(self, dtype=None, copy=None):
        return self.to_dense()


def riemannian_hessian(x, func, vector=False, metric="euclidean", check_complex=False, stiefel_elementary=True):
    """Block form of the dense Hessian, stiefel.py:477-531 with vector=True: a list over source isometries i of lists over
    target isometries j of (dof_j, dof_i) blocks -- what `riemannian_hessian_vec` stacks.  All columns come from one batched
    device evaluation (`RiemannianHessianOperator.to_dense`)."""
    if not vector:
        raise NotImplementedError("only vector=True (coefficient blocks) is supported; the reference marks the other layout for deprecation")
    metric_id(metric)
    if not stiefel_elementary:
        raise NotImplementedError("only the Stiefel elementary tangent basis is supported")
    op = RiemannianHessianOperator(x, _need_fit(func), metric)
    H = op.to_dense()
    sizes = [stiefel_tangent_dof(a) for a in op.x]
    off = np.concatenate([[0], np.cumsum(sizes)])
    return [[H[off[j]:off[j + 1], off[i]:off[i + 1]] for j in range(len(sizes))] for i in range(len(sizes))]


def riemannian_hessian_vec(x, func, transpose=False, metric="euclidean", stiefel_elementary=True):
    """Riemannian Hessian in the (A-upper, B) coefficient basis, stiefel.py:534-544, as a matrix-free operator."""
    metric_id(metric)
    if not stiefel_elementary:
        raise NotImplementedError("only the Stiefel elementary tangent basis is supported")
    op = RiemannianHessianOperator(x, _need_fit(func), metric)
    if transpose:
        return op.to_dense().T
    return op
