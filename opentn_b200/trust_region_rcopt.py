"""Mirror of `opentn.trust_region_rcopt` (same signatures).

`riemannian_trust_region_optimize(f, retract, gradfunc, hessfunc, x_init, ...)` keeps the reference's host control flow
(trust_region_rcopt.py:5-89) for arbitrary callables -- with this package's `StiefelFit`-backed callables every `f`,
`gradfunc`, `hess @ d` and `retract` call runs on the GPU.  The batched, fully device-resident solver (all instances in
lock step, no per-iteration host arithmetic) is `StiefelFit.tr_run` / `otn_tr_run`.
"""
from __future__ import annotations

import warnings

import numpy as np


def riemannian_trust_region_optimize(f, retract, gradfunc, hessfunc, x_init, save_x=False, check_convergence: bool = False,
                                     show_quotient: bool = True, **kwargs):
    """Riemannian trust region (Absil et al., Alg. 10).  Returns (x_iter, f_iter, radius) like the reference:
    f_iter[k] is the cost before step k, x_iter[-1] the iterate after the last step."""
    rho_trust = kwargs.get("rho_trust", 0.125)
    radius_init = kwargs.get("radius_init", 0.01)
    maxradius = kwargs.get("maxradius", 0.1)
    niter = kwargs.get("niter", 20)
    gfunc = kwargs.get("gfunc", None)
    tol = kwargs.get("tol", 1e-10)
    verbose = kwargs.get("verbose", True)
    tcg_kwargs = {k: kwargs["tcg_" + k] for k in ("maxiter", "abstol", "reltol") if ("tcg_" + k) in kwargs}
    assert 0 <= rho_trust < 0.25
    x = x_init
    radius = radius_init
    f_iter, g_iter, x_iter = [], [], [x]
    if gfunc is not None:
        g_iter.append(gfunc(x))
    k = 0
    try:
        for k in range(niter):
            if verbose:
                print(f"iteration: {k}")
            grad = gradfunc(x)
            hess = hessfunc(x)
            eta, on_boundary = truncated_cg(grad, hess, radius, **tcg_kwargs)
            x_next = retract(x, eta)
            fx = f(x)
            f_iter.append(fx)
            if verbose:
                print(f"f(x{k}): {fx}")
            rho = (f(x_next) - fx) / (np.dot(grad, eta) + 0.5 * np.dot(eta, hess @ eta))
            if rho < 0.25:
                radius *= 0.25
            elif rho > 0.75 and on_boundary:
                radius = min(2 * radius, maxradius)
            if show_quotient and verbose:
                print("rho:", rho, "updated radius:", radius)
            if rho > rho_trust:
                x = x_next
            if gfunc is not None:
                g_iter.append(gfunc(x))
            if save_x or k == (niter - 1):
                x_iter.append(x)
            if check_convergence and len(f_iter) > 1:
                if np.abs(f_iter[-1] - f_iter[-2]) / f_iter[-1] <= tol:
                    if verbose:
                        print(f"optimization converged prematurely at iteration: {k}")
                    break
        return x_iter, f_iter, radius
    except KeyboardInterrupt:
        print(f"optimization was stopped prematurely at iteration: {k}")
        return x_iter, f_iter, radius


def truncated_cg(grad, hess, radius, **kwargs):
    """Steihaug truncated CG (Absil et al., Alg. 11) for  min <grad,z> + <z,Hz>/2  s.t. |z| <= radius.
    `hess` only needs `@` with a 1-D vector.  trust_region_rcopt.py:92-134."""
    maxiter = kwargs.get("maxiter", 2 * len(grad))
    abstol = kwargs.get("abstol", 1e-8)
    reltol = kwargs.get("reltol", 1e-6)
    r = np.array(grad, dtype=np.float64, copy=True)
    rsq = np.dot(r, r)
    stoptol = max(abstol, reltol * np.sqrt(rsq))
    z = np.zeros_like(r)
    d = -r
    for _ in range(maxiter):
        Hd = hess @ d
        dHd = np.dot(d, Hd)
        t = _move_to_boundary(z, d, radius)
        with np.errstate(divide="ignore", invalid="ignore"):
            alpha = rsq / dHd
        if dHd <= 0 or alpha > t:
            return z + t * d, True
        r += alpha * Hd
        z += alpha * d
        rsq_next = np.dot(r, r)
        if np.sqrt(rsq_next) <= stoptol:
            return z, False
        d = -r + (rsq_next / rsq) * d
        rsq = rsq_next
    return z, False


def _move_to_boundary(b, d, radius):
    """Positive t with |b + t d| = radius, trust_region_rcopt.py:137-152."""
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


def solve_quadratic_equation(p, q):
    """Both roots of x^2 + 2 p x + q = 0, ascending; trust_region_rcopt.py:155-166."""
    disc = p ** 2 - q
    if disc < 0:
        raise ValueError("require non-negative discriminant")
    if p == 0:
        root = np.sqrt(-q)
        return (-root, root)
    x1 = -(p + np.sign(p) * np.sqrt(disc))
    return tuple(sorted((x1, q / x1)))
