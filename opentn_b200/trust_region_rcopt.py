"""Host face of the trust-region optimiser with the call signatures of `opentn.trust_region_rcopt`.

Two ways through `riemannian_trust_region_optimize`:

* **device path** -- when `f`, `retract`, `gradfunc`, `hessfunc` are the four closures handed out by
  `StiefelFit.closures(metric)` (they carry the fit they belong to), nothing of the iteration runs on the host: every
  trust-region step is one `otn_tr_run(niter=1)` call warm-started with the current radius (include/otn.h), and this
  module only keeps the reference's book-keeping (history lists, printing, premature-convergence test, Ctrl-C).
* **callable path** -- arbitrary callables (for instance the lambdas of the reference's notebooks wrapped around
  `gradient_stiefel_vec` / `riemannian_hessian_vec` / `retract_stiefel`, each of which runs on the GPU) are driven by a
  host loop organised as: `_Options` (keyword parsing) -> `_SubproblemSolver` (Steihaug truncated CG on a Hessian
  *operator*, only `hess @ v` is ever applied) -> `_RadiusPolicy` (acceptance ratio, radius update, accept / reject).

Behaviour followed (not copied) from the reference: defaults and keyword names of trust_region_rcopt.py:36-46 and
:107-109, the meaning of the returned triple `(x_iter, f_iter, radius)` (:86: `f_iter[k]` is the cost before step k,
`x_iter[-1]` the iterate after the last step), the radius rules of :68-73, the acceptance rule of :76, the premature
convergence test of :82-85 and the interrupt behaviour of :87-89; tCG exits of :117-133.
"""
from __future__ import annotations

import math
import warnings
from dataclasses import dataclass, field
from typing import Any, Callable, List, Optional, Tuple

import numpy as np

__all__ = ["riemannian_trust_region_optimize", "truncated_cg", "solve_quadratic_equation"]


# ---------------------------------------------------------------------------------------------------------------------
# options
# ---------------------------------------------------------------------------------------------------------------------
@dataclass
class _Options:
    rho_trust: float = 0.125
    radius_init: float = 0.01
    maxradius: float = 0.1
    niter: int = 20
    gfunc: Optional[Callable] = None
    tol: float = 1e-10
    verbose: bool = True
    tcg: dict = field(default_factory=dict)

    _PLAIN = ("rho_trust", "radius_init", "maxradius", "niter", "gfunc", "tol", "verbose")
    _TCG = ("maxiter", "abstol", "reltol")

    @classmethod
    def parse(cls, kwargs: dict) -> "_Options":
        opt = cls(**{name: kwargs[name] for name in cls._PLAIN if name in kwargs})
        opt.tcg = {name: kwargs["tcg_" + name] for name in cls._TCG if "tcg_" + name in kwargs}
        if not (0 <= opt.rho_trust < 0.25):
            raise AssertionError("rho_trust must lie in [0, 0.25)")
        return opt


# ---------------------------------------------------------------------------------------------------------------------
# scalar helpers of the boundary step
# ---------------------------------------------------------------------------------------------------------------------
def solve_quadratic_equation(p, q):
    """Real roots (ascending) of  t^2 + 2 p t + q = 0, computed without cancellation: the root whose two terms have the same
    sign comes from the formula, its partner from Vieta's product t1 t2 = q.  Raises ValueError on a negative discriminant
    (error contract of trust_region_rcopt.py:155-166)."""
    disc = p * p - q
    if disc < 0:
        raise ValueError("require non-negative discriminant")
    s = np.sqrt(disc)
    if p == 0:
        return (-s, s)
    far = -(p + np.sign(p) * s)          # |far| >= |near|, no cancellation
    near = q / far
    return (near, far) if near <= far else (far, near)


def _move_to_boundary(b, d, radius):
    """Step length t >= 0 with |b + t d| = radius for a point b inside the ball (the larger root of the quadratic in t).
    A zero direction cannot reach the boundary: warn and hand `b` back (what trust_region_rcopt.py:143-146 does)."""
    dd = np.dot(d, d)
    if dd == 0:
        warnings.warn("input vector 'd' is zero")
        return b
    half_lin = np.dot(b, d) / dd
    const = (np.dot(b, b) - radius ** 2) / dd
    t = solve_quadratic_equation(half_lin, const)[1]
    if t < 0:
        warnings.warn("encountered t < 0")
    return t


# ---------------------------------------------------------------------------------------------------------------------
# trust-region subproblem
# ---------------------------------------------------------------------------------------------------------------------
class _SubproblemSolver:
    """Steihaug-Toint truncated CG for  min_z <g, z> + <z, H z>/2,  |z| <= radius  (Absil et al., Alg. 11).

    `H` is anything with `H @ vector`; products are counted in `n_products`.  State lives on the object so a caller (or a
    test) can inspect the exit reason afterwards: 'negative-curvature', 'boundary', 'converged' or 'maxiter'."""

    def __init__(self, g, H, radius, maxiter=None, abstol=1e-8, reltol=1e-6):
        self.H, self.radius = H, radius
        self.residual = np.array(g, dtype=np.float64, copy=True)
        self.maxiter = 2 * len(self.residual) if maxiter is None else maxiter
        self.rr = float(np.dot(self.residual, self.residual))
        self.threshold = max(abstol, reltol * math.sqrt(self.rr))
        self.iterate = np.zeros_like(self.residual)
        self.direction = -self.residual
        self.n_products = 0
        self.exit = "maxiter"

    def _advance(self) -> Optional[Tuple[np.ndarray, bool]]:
        """One CG iteration; returns (solution, on_boundary) when the method stops, else None."""
        d = self.direction
        Hd = self.H @ d
        self.n_products += 1
        curvature = np.dot(d, Hd)
        to_edge = _move_to_boundary(self.iterate, d, self.radius)
        with np.errstate(divide="ignore", invalid="ignore"):
            step = self.rr / curvature
        if curvature <= 0 or step > to_edge:
            self.exit = "negative-curvature" if curvature <= 0 else "boundary"
            return self.iterate + to_edge * d, True
        self.residual += step * Hd
        self.iterate += step * d
        rr_new = np.dot(self.residual, self.residual)
        if np.sqrt(rr_new) <= self.threshold:
            self.exit = "converged"
            return self.iterate, False
        self.direction = -self.residual + (rr_new / self.rr) * d
        self.rr = rr_new
        return None

    def solve(self) -> Tuple[np.ndarray, bool]:
        for _ in range(self.maxiter):
            done = self._advance()
            if done is not None:
                return done
        return self.iterate, False


def truncated_cg(grad, hess, radius, **kwargs):
    """`(z, on_boundary)` of the trust-region subproblem; keywords `maxiter` (default 2 len(grad)), `abstol` (1e-8),
    `reltol` (1e-6).  `hess` only needs `@` with a 1-D vector (trust_region_rcopt.py:92-134)."""
    return _SubproblemSolver(grad, hess, radius, kwargs.get("maxiter"), kwargs.get("abstol", 1e-8),
                             kwargs.get("reltol", 1e-6)).solve()


# ---------------------------------------------------------------------------------------------------------------------
# acceptance / radius policy
# ---------------------------------------------------------------------------------------------------------------------
@dataclass
class _RadiusPolicy:
    rho_trust: float
    maxradius: float
    shrink_below: float = 0.25
    grow_above: float = 0.75

    def update(self, radius: float, rho: float, on_boundary: bool) -> float:
        if rho < self.shrink_below:
            return radius * 0.25
        if rho > self.grow_above and on_boundary:
            return min(2 * radius, self.maxradius)
        return radius

    def accepts(self, rho: float) -> bool:
        return bool(rho > self.rho_trust)


# ---------------------------------------------------------------------------------------------------------------------
# one step, two implementations
# ---------------------------------------------------------------------------------------------------------------------
@dataclass
class _StepResult:
    fx: Any               # cost at the point the step started from
    rho: Any
    radius: float
    x: Any                # iterate after accept / reject


def _step_callables(f, retract, gradfunc, hessfunc, x, radius, policy: _RadiusPolicy, tcg: dict) -> _StepResult:
    g = gradfunc(x)
    H = hessfunc(x)
    eta, on_boundary = truncated_cg(g, H, radius, **tcg)
    trial = retract(x, eta)
    fx = f(x)
    predicted = np.dot(g, eta) + 0.5 * np.dot(eta, H @ eta)
    rho = (f(trial) - fx) / predicted
    return _StepResult(fx, rho, policy.update(radius, rho, on_boundary), trial if policy.accepts(rho) else x)


def _step_device(fit, x, radius, opt: _Options) -> _StepResult:
    xs = np.stack([np.asarray(a, dtype=np.float64) for a in x])
    x_new, rep = fit.tr_run(xs, niter=1, rho_trust=opt.rho_trust, maxradius=opt.maxradius, radius=radius,
                            tcg_maxiter=opt.tcg.get("maxiter", 0), tcg_abstol=opt.tcg.get("abstol", 1e-8),
                            tcg_reltol=opt.tcg.get("reltol", 1e-6))
    return _StepResult(float(rep["f_iter"][0, 0]), float(rep["rho"][0, 0]), float(rep["radius"][0]), list(x_new))


def _shared_fit(*callables):
    """The StiefelFit all four closures were made by (`StiefelFit.closures`), or None."""
    owners = [getattr(c, "_otn_fit", None) for c in callables]
    first = owners[0]
    if first is None or any(o is None or o[0] is not first[0] or o[1] != first[1] for o in owners):
        return None
    fit, metric = first
    from .fit import metric_alphas
    if metric_alphas(fit.metric) != metric_alphas(metric):
        fit.set_metric(metric)
    return fit


def riemannian_trust_region_optimize(f, retract, gradfunc, hessfunc, x_init, save_x=False, check_convergence: bool = False,
                                     show_quotient: bool = True, **kwargs):
    """Riemannian trust region (Absil et al., Alg. 10) with the reference's signature and return value
    `(x_iter, f_iter, radius)`; extra keyword `verbose=False` silences the per-iteration prints."""
    opt = _Options.parse(kwargs)
    policy = _RadiusPolicy(opt.rho_trust, opt.maxradius)
    fit = _shared_fit(f, retract, gradfunc, hessfunc)
    say = print if opt.verbose else (lambda *a, **k: None)

    x, radius = x_init, opt.radius_init
    xs_hist: List[Any] = [x]
    f_hist: List[Any] = []
    g_hist: List[Any] = [] if opt.gfunc is None else [opt.gfunc(x)]
    k = 0
    try:
        while k < opt.niter:
            say(f"iteration: {k}")
            res = _step_device(fit, x, radius, opt) if fit is not None else \
                _step_callables(f, retract, gradfunc, hessfunc, x, radius, policy, opt.tcg)
            f_hist.append(res.fx)
            say(f"f(x{k}): {res.fx}")
            x, radius = res.x, res.radius
            if show_quotient:
                say("rho:", res.rho, "updated radius:", radius)
            if opt.gfunc is not None:
                g_hist.append(opt.gfunc(x))
            last = k == opt.niter - 1
            if save_x or last:
                xs_hist.append(x)
            if check_convergence and len(f_hist) > 1 and abs(f_hist[-1] - f_hist[-2]) / f_hist[-1] <= opt.tol:
                say(f"optimization converged prematurely at iteration: {k}")
                break
            k += 1
    except KeyboardInterrupt:
        print(f"optimization was stopped prematurely at iteration: {k}")
    return xs_hist, f_hist, radius
