"""Host-side mirror of `opentn.optimization` for the Stiefel trust-region path.

`model_stiefel_local`, `model_Ys_stiefel` and `frobenius_norm` keep the reference's signatures
(optimization.py:106, 147, 56) and evaluate on the GPU through libotn.  The cost closure the reference notebooks build
from them is `opentn_b200.StiefelFit` here (see fit.py).
"""
from __future__ import annotations

import numpy as np

from .fit import StiefelFit
from .transformations import (choi2ortho, create_2local_liouvillians, exp_operator_dt, factorize_psd_truncated,
                              get_kitaev_nn_linbladian, lindbladian2super, super2choi)

_MODEL_CACHE = {}


def _model_fit(model, N, d, D, device=0):
    key = (model, N, d, device)
    fit = _MODEL_CACHE.get(key)
    if fit is None:
        fit = StiefelFit(np.zeros((D, D)), model=model, N=N, d=d, device=device)
        _MODEL_CACHE[key] = fit
    return fit


def model_stiefel_local(xs, N: int, d: int) -> np.ndarray:
    """Composition of m two-site layers (odd list index => pbc-shifted pairs), optimization.py:106-137."""
    return _model_fit("local", N, d, d ** (2 * N)).model_matrix(xs)


def model_Ys_stiefel(xs) -> np.ndarray:
    """Composition of m full-sites layers, optimization.py:147-151.  d = 2 is assumed to infer N from the isometry width."""
    p = np.asarray(xs[0]).shape[-1]
    N = int(round(np.log2(p)))
    if 2 ** N != p:
        raise ValueError("model_Ys_stiefel: isometry width must be 2^N")
    return _model_fit("full", N, 2, p * p).model_matrix(xs)


def frobenius_norm(A, B, squared: bool = False):
    """|| A - B ||_F (squared if asked), optimization.py:56-61: A real (a model matrix), B real or complex (a target).  GPU
    reduction; inside a fit the residual norm is fused into the last chain GEMM instead (StiefelFit.cost)."""
    from . import _lib
    from . import transformations as _T
    A = np.asarray(A)
    B = np.asarray(B)
    if np.iscomplexobj(A) and np.abs(A.imag).max() > 0 and not (np.iscomplexobj(B) and np.abs(B.imag).max() > 0):
        A, B = B, A                                     # the norm is symmetric: compute_loss passes (exact, prediction)
    if np.iscomplexobj(A):
        if np.abs(A.imag).max() > 0:
            raise NotImplementedError("frobenius_norm: one argument (the model matrix) must be real on the accelerated path")
        A = A.real
    a = _lib.as_f64(A).reshape(-1)
    bre = _lib.as_f64(B.real).reshape(-1)
    bim = _lib.as_f64(B.imag).reshape(-1) if np.iscomplexobj(B) else None
    if a.shape != bre.shape:
        raise ValueError("operands could not be broadcast together")
    out = np.empty(1)
    _lib.check(_lib.load().otn_frobenius(_lib.ptr(a), _lib.ptr(bre), _lib.ptr(bim), 1, a.size, int(bool(squared)), _lib.ptr(out),
                                         _T._DEVICE))
    return float(out[0])


def compute_loss(xi, loss_fn, model, exact, **kwargs):
    """loss_fn(exact, model(xi, **kwargs)), optimization.py:153-159 (the reference wraps loss_fn in `jit`; here `model` and
    `loss_fn` are the device-backed functions of this module)."""
    prediction = model(xi, **kwargs)
    exact = np.asarray(exact)
    assert exact.ndim == np.ndim(prediction) == 2, "not a matrix"
    return loss_fn(exact, prediction)


def get_general_trotter_local_ansatz(lindbladians, tau, n: int):
    """[half] + [full]*(2n-1) + [half] two-site superoperators of the Strang splitting, optimization.py:172-185."""
    if not isinstance(lindbladians, list):
        lindbladians = [lindbladians]
    dt = tau / n
    s = lindbladian2super(Li=lindbladians)
    half, full = exp_operator_dt(s, dt / 2), exp_operator_dt(s, dt)
    return [half] + [full] * (2 * n - 1) + [half]


def get_kitaev_trotter_local_ansatz(gamma: float, tau, n: int):
    """optimization.py:161-169"""
    return get_general_trotter_local_ansatz([get_kitaev_nn_linbladian(gamma)], tau, n)


def compute_trotter_approximation_error(d: int, N: int, tau, n: int, Li) -> float:
    """Cost of the Strang ansatz at its natural Choi rank, optimization.py:215-226 (cost evaluated on the GPU)."""
    Lvec = create_2local_liouvillians(Li=Li, N=N, d=d, pbc=True)[0]
    chois = [super2choi(x.real) for x in get_general_trotter_local_ansatz(Li, tau, n)]
    ranks = [np.linalg.matrix_rank(c) for c in chois]
    assert len(set(ranks)) == 1, "not all ranks are the same"
    xs = [choi2ortho(factorize_psd_truncated(c, chi_max=ranks[0])) for c in chois]
    fit = StiefelFit(exp_operator_dt(Lvec, tau), model="local", N=N, d=d)
    try:
        return fit(xs)
    finally:
        fit.close()


def compute_kitaev_approximation_error(d: int, N: int, gamma: float, tau, n: int) -> float:
    """optimization.py:187-213"""
    return compute_trotter_approximation_error(d, N, tau, n, [get_kitaev_nn_linbladian(gamma)])
