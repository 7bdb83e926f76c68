"""Host-side mirror of `opentn.transformations` for the Stiefel trust-region path (same names, argument meaning, errors).

Two kinds of functions live here:

* hot-path arithmetic (sites A, B, C of SURVEY 2.1) -- `ortho2super`, `kraus2superop`, `create_supertensored_from_local`
  + `convert_supertensored2liouvillianfull` (fused as `local2full`), `compose_superops_list` -- run on the GPU through
  libotn (include/otn.h); they raise `OtnError` when the library or a GPU is missing.
* problem set-up that the reference runs once per instance family (Liouvillian construction, `expm`, truncated PSD
  factorisation for the starting point; SURVEY 2 rows 5-6, "out of scope as a kernel") -- plain NumPy/SciPy on the host,
  written against the reference's conventions (row-major vectorisation, index layouts of SURVEY appendix B).
  Pure index reshuffles (`ortho2choi`, `choi2ortho`, `super2choi`, `choi2super`) are NumPy views.
"""
from __future__ import annotations

from math import log

import numpy as np
import scipy.linalg

from . import _lib
from ._lib import as_f64, check, ptr

_DEVICE = 0

# ---------------------------------------------------------------------------------------------------------------------
# set-up layer (host)
# ---------------------------------------------------------------------------------------------------------------------
_I2 = np.eye(2, dtype=complex)
X = np.array([[0, 1], [1, 0]], dtype=complex)
Y = np.array([[0, -1j], [1j, 0]], dtype=complex)
Z = np.array([[1, 0], [0, -1]], dtype=complex)
I = _I2  # noqa: E741  (the reference exports this name, opentn/states/qubits.py:13)


def get_ladder_operator(num_levels: int = 2, adjoint: bool = False) -> np.ndarray:
    """Annihilation (or creation) operator, opentn/states/qubits.py:19-37."""
    a = np.zeros((num_levels, num_levels), dtype=complex)
    for i in range(1, num_levels):
        a[i - 1, i] = np.sqrt(i)
    return a.conj().T if adjoint else a


def exp_operator_dt(op: np.ndarray, tau: float = 1, backend: str = "jax") -> np.ndarray:
    """exp(tau * op), transformations.py:24-32.  backend 'scipy' / 'jax' (the reference's two names; both are Pade scaling-and-
    squaring, evaluated with SciPy on the host) or 'cuda' (otn_expm: Taylor scaling-and-squaring on the FP64 tensor pipe; the
    4096 x 4096 Liouvillian of N = 6 takes well under a second instead of ~15-90 s)."""
    if backend == "cuda":
        a = np.asarray(op)
        D = a.shape[0]
        assert a.ndim == 2 and a.shape[1] == D, "square matrix expected"
        lib = _lib.load()
        re = as_f64(a.real)
        out_re = np.empty((D, D))
        if np.iscomplexobj(a) and np.any(a.imag):
            im = as_f64(a.imag)
            out_im = np.empty((D, D))
            check(lib.otn_expm(ptr(re), ptr(im), D, float(tau), ptr(out_re), ptr(out_im), _DEVICE))
            return out_re + 1j * out_im
        check(lib.otn_expm(ptr(re), None, D, float(tau), ptr(out_re), None, _DEVICE))
        return out_re.astype(a.dtype) if np.iscomplexobj(a) else out_re
    if backend not in ("scipy", "jax"):
        raise ValueError(f"{backend} is not a valid backend string")
    return scipy.linalg.expm(np.asarray(op) * tau)


def exp_operator_sweep(op: np.ndarray, taus) -> np.ndarray:
    """[exp(tau * op) for tau in taus] as one array (len(taus), D, D): the tau sweep of one generator in one batched sequence of
    launches (otn_expm_batch).  No counterpart in the reference, which loops `exp_operator_dt` (transformations.py:24-32) over
    tau on the host; used by `create_trotter_layers_sweep`."""
    a = np.asarray(op)
    D = a.shape[0]
    assert a.ndim == 2 and a.shape[1] == D, "square matrix expected"
    t = as_f64(np.atleast_1d(np.asarray(taus, dtype=np.float64)))
    lib = _lib.load()
    re = as_f64(a.real)
    out_re = np.empty((t.size, D, D))
    if np.iscomplexobj(a) and np.any(a.imag):
        im = as_f64(a.imag)
        out_im = np.empty((t.size, D, D))
        check(lib.otn_expm_batch(ptr(re), ptr(im), D, ptr(t), t.size, ptr(out_re), ptr(out_im), _DEVICE))
        return out_re + 1j * out_im
    check(lib.otn_expm_batch(ptr(re), None, D, ptr(t), t.size, ptr(out_re), None, _DEVICE))
    return out_re.astype(a.dtype) if np.iscomplexobj(a) else out_re


def create_trotter_layers_sweep(liouvillians: list, taus, order: int = 2, half_idx: int = 1) -> list:
    """`create_trotter_layers` (transformations.py:387-402) for every tau of a sweep: returns one list per tau, each
    [exp(tau L), exp(tau/2 L_odd), exp(tau L_even)]; three batched device `expm` sweeps in total."""
    if order != 2:
        raise ValueError("ATM only order 2 is implemented")
    taus = np.atleast_1d(np.asarray(taus, dtype=np.float64))
    per_op = [exp_operator_sweep(op, taus / 2 if i == half_idx else taus) for i, op in enumerate(liouvillians)]
    return [[per_op[i][k] for i in range(len(liouvillians))] for k in range(taus.size)]


def _liouvillian_terms_device(ops, terms, N: int, d: int) -> np.ndarray:
    """sum of dissipators of two-site jump operators placed on site pairs (otn_liouvillian).  terms: (op index, site0, site1)."""
    ops = np.stack([np.asarray(o, dtype=complex) for o in ops])
    q = d * d
    assert ops.shape[1:] == (q, q), "two-site jump operators expected"
    D = d ** (2 * N)
    re, im = as_f64(ops.real), as_f64(ops.imag)
    t = np.asarray(terms, dtype=np.int32).reshape(-1, 3)
    t_op, t_s0, t_s1 = (np.ascontiguousarray(t[:, i]) for i in range(3))
    out_re, out_im = np.empty((D, D)), np.empty((D, D))
    lib = _lib.load()
    check(lib.otn_liouvillian(ptr(re), ptr(im), ops.shape[0], N, d, ptr(t_op), ptr(t_s0), ptr(t_s1), t.shape[0], ptr(out_re),
                              ptr(out_im), _DEVICE))
    return out_re + 1j * out_im


def vectorize(matrix: np.ndarray) -> np.ndarray:
    """Row-major vectorisation to a column vector, transformations.py:221-223."""
    return np.reshape(matrix, (-1, 1))


def unvectorize(vector: np.ndarray, d0=None, d1=None) -> np.ndarray:
    """transformations.py:225-234"""
    size = np.size(vector)
    if not d0 and not d1:
        d0 = d1 = int(np.sqrt(size))
    elif not d0:
        d0 = size // d1
    elif not d1:
        d1 = size // d0
    return np.reshape(vector, (d0, d1))


def vectorize_hamiltonian(H: np.ndarray, dim: int = None) -> np.ndarray:
    """H rho - rho H -> (H (x) 1 - 1 (x) H^T)|rho>>, transformations.py:236-241 (the i is included in H)."""
    dim = dim or H.shape[0]
    eye = np.eye(dim, dtype=complex)
    return np.kron(H, eye) - np.kron(eye, H.T)


def vectorize_dissipative(L: np.ndarray, dim: int = None) -> np.ndarray:
    """L rho L^+ - {L^+ L, rho}/2 in row-major vectorisation, transformations.py:243-250."""
    dim = dim or L.shape[0]
    eye = np.eye(dim, dtype=complex)
    LdL = L.conj().T @ L
    return np.kron(L, L.conj()) - 0.5 * np.kron(LdL, eye) - 0.5 * np.kron(eye, LdL.T)


def lindbladian2super(H: np.ndarray = None, Li: list = [], dim: int = None) -> np.ndarray:
    """Liouvillian superoperator of a Lindblad master equation (no exponential), transformations.py:34-66."""
    if not dim:
        dim = H.shape[0] if H is not None else Li[0].shape[0]
    out = np.zeros((dim * dim, dim * dim), dtype=complex)
    if H is not None:
        out += vectorize_hamiltonian(H, dim)
    for L in Li:
        out += vectorize_dissipative(L, dim)
    return out


def op2fullspace(op: np.ndarray, i: int, N: int, num_sites: int = 1):
    """Embed a local operator acting on sites i..i+num_sites-1 into N qubits, transformations.py:252-260."""
    assert 0 <= i < N, "i needs to be greater or equal than 0 and smaller than N!"
    assert N > 0, "N needs to be greater than 0!"
    if N == 1:
        return op
    return np.kron(np.eye(2 ** i), np.kron(op, np.eye(2 ** (N - i - num_sites))))


def dissipative2liouvillian_full(L: np.ndarray, i: int, N: int, num_sites: int = 1):
    """transformations.py:262-272"""
    return vectorize_dissipative(op2fullspace(L, i, N, num_sites))


def permute_cyclic(a: list, n: int = 1, direction: str = "left") -> list:
    """transformations.py:446-453"""
    if direction == "left":
        return a[n:] + a[:n]
    if direction == "right":
        return a[-n:] + a[:-n]
    raise ValueError(direction)


def permute_operator(op: np.ndarray, permutation: list, d: int):
    """Relabel lattice-site legs of an operator, transformations.py:455-464."""
    N = len(permutation)
    assert op.shape == (d ** N, d ** N)
    t = np.reshape(op, [d] * (2 * N)).transpose(list(permutation) + [N + s for s in permutation])
    return np.reshape(t, (d ** N, d ** N))


def permute_operator_pbc(op: np.ndarray, N: int, d: int, direction: str = "right"):
    """transformations.py:466-471"""
    return permute_operator(op, permute_cyclic(list(range(N)), direction=direction), d)


def create_pbc_liouvillian(Lnn: np.ndarray, N: int, d: int):
    """Dissipator of the pair (N-1, 0), transformations.py:364-369."""
    return vectorize_dissipative(permute_operator_pbc(op2fullspace(Lnn, N - 2, N, 2), N, d))


def local_lindbladian_to_full_liouvillians(Lnn: np.ndarray, N: int, d: int, pbc: bool = False):
    """(total, odd, even) Liouvillians from a two-site jump operator, transformations.py:337-362."""
    num_sites = int(log(Lnn.shape[0], d))
    assert num_sites <= N, f"the numbert of sites on the lindbladians: {num_sites} should be smaller or equal than N: {N} "
    if N == num_sites or N == 1:
        pbc = False
    D = d ** (2 * N)
    odd = np.zeros((D, D), dtype=complex)
    even = np.zeros((D, D), dtype=complex)
    for i in range(0, N, 2):
        odd += dissipative2liouvillian_full(Lnn, i, N, num_sites)
    for i in range(1, N - 1, 2):
        even += dissipative2liouvillian_full(Lnn, i, N, num_sites)
    if pbc:
        even += create_pbc_liouvillian(Lnn, N, d)
    return even + odd, odd, even


def create_2local_liouvillians(Li, N: int, d: int, pbc: bool = False, backend: str = "numpy"):
    """transformations.py:324-335.  backend='cuda' assembles the odd and even sums entry by entry on the device
    (otn_liouvillian) instead of through D x D Kronecker products on the host."""
    if not isinstance(Li, list):
        Li = [Li]
    if backend == "cuda":
        if N % 2:
            raise ValueError("two-site odd/even layers need an even number of sites (the reference's loop "
                             "`range(0, N, 2)`, transformations.py:351, runs off the chain for odd N)")
        use_pbc = pbc and N > 2                         # transformations.py:347-348: pbc is dropped when N == num_sites
        odd_t = [(j, i, i + 1) for j in range(len(Li)) for i in range(0, N - 1, 2)]
        even_t = [(j, i, i + 1) for j in range(len(Li)) for i in range(1, N - 1, 2)]
        if use_pbc:
            even_t += [(j, N - 1, 0) for j in range(len(Li))]   # permute_operator_pbc(direction='right'): first factor on N-1
        D = d ** (2 * N)
        odd = _liouvillian_terms_device(Li, odd_t, N, d)
        even = _liouvillian_terms_device(Li, even_t, N, d) if even_t else np.zeros((D, D), dtype=complex)
        return odd + even, odd, even
    if backend != "numpy":
        raise ValueError(f"{backend} is not a valid backend string")
    D = d ** (2 * N)
    odd = np.zeros((D, D), dtype=complex)
    even = np.zeros((D, D), dtype=complex)
    for L in Li:
        _, o, e = local_lindbladian_to_full_liouvillians(L, N, d, pbc)
        odd += o
        even += e
    return odd + even, odd, even


def get_kitaev_nn_linbladian(gamma: float):
    """sqrt(gamma) (a0^+ + a1^+)(a0 - a1)/4, transformations.py:371-377."""
    lo, ra = get_ladder_operator(), get_ladder_operator(adjoint=True)
    plus = op2fullspace(ra, 0, 2) + op2fullspace(ra, 1, 2)
    minus = op2fullspace(lo, 0, 2) - op2fullspace(lo, 1, 2)
    return np.sqrt(gamma) * plus @ minus / 4


def create_kitaev_liouvillians(N: int, d: int, gamma: float, pbc: bool = False, backend: str = "numpy"):
    """transformations.py:379-384"""
    Lnn = get_kitaev_nn_linbladian(gamma)
    full, odd, even = create_2local_liouvillians(Lnn, N, d, pbc, backend=backend)
    return full, odd, even, Lnn


def create_trotter_layers(liouvillians: list, tau: float = 1, order: int = 2, half_idx: int = 1, backend: str = "jax"):
    """transformations.py:387-402"""
    if order != 2:
        raise ValueError("ATM only order 2 is implemented")
    return [exp_operator_dt(op, tau / 2 if i == half_idx else tau, backend) for i, op in enumerate(liouvillians)]


# ---- index reshuffles (views) -------------------------------------------------------------------------------------------
def super2choi(superop: np.ndarray, dim: int = None) -> np.ndarray:
    """transformations.py:68-94"""
    if not dim:
        dim = int(np.sqrt(superop.shape[0]))
    assert superop.shape[0] == dim ** 2
    return np.reshape(superop, [dim] * 4).swapaxes(1, 2).reshape(dim * dim, dim * dim)


def choi2super(choi: np.ndarray, dim: int = None) -> np.ndarray:
    """transformations.py:96-125"""
    if not dim:
        dim = int(np.sqrt(choi.shape[0]))
    assert choi.shape[0] == dim ** 2
    return np.reshape(choi, [dim] * 4).swapaxes(1, 2).reshape(dim * dim, dim * dim)


def choi2ortho(x: np.ndarray, dim: int = None) -> np.ndarray:
    """Choi factor (dim^2, K) -> isometry (dim K, dim), transformations.py:701-714."""
    assert x.ndim == 2, "x should be a matrix"
    if not dim:
        dim = int(np.sqrt(x.shape[0]))
    k = x.shape[1]
    return np.reshape(x, [dim, dim, k]).swapaxes(1, 2).reshape(dim * k, dim)


def ortho2choi(x: np.ndarray, dim: int = None) -> np.ndarray:
    """Isometry (dim K, dim) -> Choi factor (dim^2, K), transformations.py:716-726."""
    if not dim:
        dim = x.shape[1]
    k = int(x.shape[0] / dim)
    return np.reshape(x, [dim, k, dim]).swapaxes(1, 2).reshape(dim * dim, k)


# ---- starting points (host LAPACK; once per instance) -------------------------------------------------------------
def split_matrix_svd(op: np.ndarray, chi_max: int = 2, eps: float = 1e-9):
    """transformations.py:752-790"""
    assert op.ndim == 2
    u, s, v = np.linalg.svd(op, full_matrices=False)
    keep = chi_max if eps is None else min(chi_max, int(np.sum(s > eps)))
    assert keep >= 1
    idx = np.argsort(s)[::-1][:keep]
    return u[:, idx], s[idx], v[idx, :]


def factorize_psd_truncated(psd: np.ndarray, chi_max: int = 2, eps: float = None) -> np.ndarray:
    """transformations.py:793-802"""
    u, s, _ = split_matrix_svd(psd, chi_max, eps)
    return u * np.sqrt(s)[None, :]


def factorize_psd(psd: np.ndarray, check_hermitian: bool = False, tol: float = 1e-9):
    """transformations.py:284-322"""
    if check_hermitian:
        assert np.allclose(psd, psd.conj().T), "Matrix is not hermitian"
    w, v = np.linalg.eigh(psd)
    w = np.where(np.abs(w) < tol, 0.0, w)
    if np.any(w < 0):
        raise ValueError(f"invalid eigenvalue found: {w[w < 0][0]}. input is not PSD")
    return v * np.sqrt(w)[None, :]


def unfactorize_psd(x: np.ndarray) -> np.ndarray:
    """transformations.py:804-806"""
    return x @ x.conj().T


def super2ortho_batch(superops, rank: int) -> np.ndarray:
    """super2ortho (transformations.py:728-741) of a stack of real 2-site superoperators (count, 16, 16) on the device
    (otn_super2ortho: Jacobi eigen-decomposition of the Choi matrices, one warp each) -> (count, 4 rank, 4).  Kraus operators come
    out with a fixed sign convention; ortho2super(x) equals the host route's."""
    a = np.asarray(superops)
    if np.iscomplexobj(a):
        if np.abs(a.imag).max() > 0:
            raise NotImplementedError("super2ortho_batch: real superoperators only (the reference passes `.real`)")
        a = a.real
    a = as_f64(a)
    assert a.ndim == 3 and a.shape[1] == a.shape[2], "stack of square superoperators expected"
    dim = int(round(np.sqrt(a.shape[1])))
    assert dim * dim == a.shape[1]
    out = np.empty((a.shape[0], dim * rank, dim))
    check(_lib.load().otn_super2ortho(ptr(a), a.shape[0], dim, int(rank), ptr(out), _DEVICE))
    return out


def super2ortho(x: np.ndarray, rank: int = None, backend: str = "numpy") -> np.ndarray:
    """superop -> Choi -> rank-truncated factor -> isometry, transformations.py:728-741."""
    if backend == "cuda":
        if not rank:
            rank = np.linalg.matrix_rank(super2choi(np.asarray(x)))
        return super2ortho_batch(np.asarray(x)[None], rank)[0]
    c = super2choi(x)
    if not rank:
        rank = np.linalg.matrix_rank(c)
    return choi2ortho(factorize_psd_truncated(c, chi_max=rank))


# ---------------------------------------------------------------------------------------------------------------------
# hot-path arithmetic (GPU through libotn)
# ---------------------------------------------------------------------------------------------------------------------
def set_device(device: int):
    global _DEVICE
    _DEVICE = int(device)


def _real_input(a, what):
    a = np.asarray(a)
    if np.iscomplexobj(a):
        if np.abs(a.imag).max() > 0:
            raise NotImplementedError(f"{what}: complex input is not supported on the accelerated path (the reference's "
                                      "optimisation variables are real, SURVEY section 0)")
        a = a.real
    return a


def ortho2super(x: np.ndarray) -> np.ndarray:
    """sum_k E_k (x) E_k^*, E_k = x[k::K, :]; transformations.py:743-749.  x: (dim K, dim) or a stack (..., dim K, dim)."""
    x = _real_input(x, "ortho2super")
    lead = x.shape[:-2]
    n, dim = x.shape[-2:]
    if n % dim:
        raise ValueError("isometry must have shape (dim*K, dim)")
    K = n // dim
    xb = as_f64(x).reshape(-1, n, dim)
    out = np.empty((xb.shape[0], dim * dim, dim * dim))
    check(_lib.load().otn_ortho2super(ptr(xb), xb.shape[0], dim, K, ptr(out), _DEVICE))
    return out.reshape(lead + (dim * dim, dim * dim))


def kraus2superop(kraus_list) -> np.ndarray:
    """sum_k E_k (x) E_k^*, transformations.py:170-180 (`np.kron(kraus, kraus.conj())` summed over the list).  Real or complex
    square Kraus operators; returned complex like the reference.  One fused kernel (`otn_kraus2superop`): the Kronecker
    products are never formed one by one."""
    if not isinstance(kraus_list, list):
        kraus_list = [kraus_list]
    E = np.stack([np.asarray(k) for k in kraus_list])  # (K, rows, cols)
    K, rows, cols = E.shape
    if rows != cols:
        raise NotImplementedError("kraus2superop: only square Kraus operators on the accelerated path")
    cplx = np.iscomplexobj(E)
    e_re = as_f64(E.real)[None]
    e_im = as_f64(E.imag)[None] if cplx else None
    o_re = np.empty((1, rows * rows, rows * rows))
    o_im = np.empty_like(o_re) if cplx else None
    check(_lib.load().otn_kraus2superop(ptr(e_re), ptr(e_im), 1, K, rows, ptr(o_re), ptr(o_im), _DEVICE))
    return (o_re[0] + 1j * o_im[0]) if cplx else o_re[0].astype(complex)


def compose_superops_list(superops) -> np.ndarray:
    """Y_m @ ... @ Y_1 (list order = application order), transformations.py:841-851."""
    S = np.stack([_real_input(s, "compose_superops_list") for s in superops])
    m, D, D2 = S.shape
    assert D == D2
    S = as_f64(S)[None]
    out = np.empty((1, D, D))
    check(_lib.load().otn_compose(ptr(S), 1, m, D, ptr(out), _DEVICE))
    return out[0]


def local2full(localop: np.ndarray, N: int, d: int = 2, shift_pbc: bool = False) -> np.ndarray:
    """Fused `convert_supertensored2liouvillianfull(create_supertensored_from_local(localop, N, pbc=True), N, d, shift_pbc)`
    (optimization.py:127-136): never materialises the un-permuted Kronecker power."""
    y = as_f64(_real_input(localop, "local2full"))
    q = d ** 4
    lead = y.shape[:-2]
    assert y.shape[-2:] == (q, q), "local superoperator must be (d^4, d^4)"
    yb = y.reshape(-1, q, q)
    D = d ** (2 * N)
    out = np.empty((yb.shape[0], D, D))
    check(_lib.load().otn_embed_local(ptr(yb), yb.shape[0], N, d, int(bool(shift_pbc)), ptr(out), _DEVICE))
    return out.reshape(lead + (D, D))


def get_indices_supertensored2liouvillianfull(N: int):
    """transformations.py:431-444"""
    src = []
    for i in range((N - 2) // 2):
        src += [2 + 4 * i, 3 + 4 * i]
    dst = list(range(N, 2 * N - 2))
    return src + [s + 2 * N for s in src], dst + [t + 2 * N for t in dst]


def create_supertensored_from_local(localop: np.ndarray, N: int, pbc: bool = False, layer: int = 0):
    """Kronecker power of a two-site superoperator, transformations.py:405-429.  Kept for signature compatibility (host
    NumPy: a pure outer product that the accelerated models never materialise -- they call the fused `local2full`)."""
    assert N % 2 == 0, "Only even number of sites allowed"
    if layer % 2 == 0:
        pbc = True
    out = np.asarray(localop) if pbc else np.eye(np.asarray(localop).shape[0])
    for _ in range(N // 2 - 1):
        out = np.kron(out, localop)
    return out


def swap_superop_indices(superop, source_indices, destination_indices, N, d, shift_pbc=False):
    """transformations.py:473-482 (pure permutation, host view + copy)."""
    t = np.moveaxis(np.reshape(superop, [d] * (4 * N)), source_indices, destination_indices)
    if shift_pbc:
        idx = permute_cyclic(list(range(N)), 1) + permute_cyclic(list(range(N, 2 * N)), 1)
        t = np.transpose(t, idx + [i + 2 * N for i in idx])
    return np.reshape(t, np.shape(superop))


def convert_supertensored2liouvillianfull(local_tensored, N: int, d: int, shift_pbc: bool = False):
    """transformations.py:484-490"""
    src, dst = get_indices_supertensored2liouvillianfull(N)
    return swap_superop_indices(local_tensored, src, dst, N, d, shift_pbc)


def convert_liouvillianfull2supertensored(full_liouvillian, N: int, d: int):
    """transformations.py:493-498"""
    dst, src = get_indices_supertensored2liouvillianfull(N)
    return swap_superop_indices(full_liouvillian, src, dst, N, d)
