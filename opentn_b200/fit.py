"""`StiefelFit`: the cost closure of the reference notebooks as an object the accelerated path can recognise.

The reference builds, in user code (experiments/trust_region_pauli_noise.ipynb cell 21),

    f_stiefel = lambda xi: frobenius_norm(model_stiefel_local(xi, N, d), exp_Lvec)

and hands it to `gradient_stiefel_vec(xi, f_stiefel, metric)` / `riemannian_hessian_vec(xi, f_stiefel, metric)`, which
differentiate the opaque closure with JAX.  Here the same closure is `StiefelFit(exp_Lvec, model='local', N=N, d=d)`:
callable as `f(xi) -> float`, and carrying batched `.cost / .cost_grad / .triple / .hvp / .tr_run` methods that run
on the GPU through libotn (include/otn.h).  Arbitrary Python closures are not supported (no CPU fallback).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import MODEL_FULL, MODEL_LOCAL, METRIC_CANONICAL, METRIC_EUCLIDEAN, OtnError, as_f64, check, ptr

_METRICS = {"euclidean": METRIC_EUCLIDEAN, "canonical": METRIC_CANONICAL}


def metric_id(metric) -> int:
    """Preset id of a metric name; -1 for a general (alpha0, alpha1) pair (`metric_alphas` validates it)."""
    if isinstance(metric, str):
        if metric not in _METRICS:
            raise ValueError(f"{metric} is not a valid metric value")  # same message as stiefel.py:461
        return _METRICS[metric]
    metric_alphas(metric)
    return -1


def metric_alphas(metric):
    """(alpha0, alpha1) of a metric: 'euclidean' = (1, 1), 'canonical' = (1, 1/2) (stiefel.py:456-459), or any positive pair
    of the family `gradient_stiefel_general` takes (stiefel.py:411-426)."""
    if isinstance(metric, str):
        return (1.0, 1.0) if metric_id(metric) == METRIC_EUCLIDEAN else (1.0, 0.5)
    try:
        a0, a1 = (float(v) for v in metric)
    except (TypeError, ValueError):
        raise ValueError(f"{metric} is not a valid metric value") from None
    if not (a0 > 0 and a1 > 0 and np.isfinite(a0) and np.isfinite(a1)):
        raise ValueError("alpha0 and alpha1 must be positive and finite")
    return a0, a1


def _is_cuda_tensor(a) -> bool:
    return hasattr(a, "is_cuda") and bool(a.is_cuda)


def _promote_pair(x, e):
    """Isometries and directions of one call must agree on real vs complex: a real array next to a complex one is promoted."""
    cx, ce = _is_cplx(x), _is_cplx(e)
    if cx == ce:
        return x, e
    def up(a):
        if isinstance(a, np.ndarray):
            return np.ascontiguousarray(a, dtype=np.complex128)
        import torch
        return a.to(torch.complex128).contiguous()
    return (x, up(e)) if cx else (up(x), e)


def _is_cplx(a) -> bool:
    if isinstance(a, np.ndarray):
        return np.iscomplexobj(a)
    return bool(getattr(a, "is_complex", lambda: False)())


class _Handle:
    """Owns one otn_problem."""

    def __init__(self, N, d, m, K, model, metric, max_batch, targets_re, targets_im, device, flags=0):
        lib = _lib.load()
        self._lib = lib
        self.h = C.c_void_p()
        n_targets = targets_re.shape[0]
        check(lib.otn_create_ex(C.byref(self.h), N, d, m, K, model, metric, max_batch, n_targets, ptr(targets_re),
                                ptr(targets_im), device, flags))
        n, p, D, mm = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        ws = C.c_longlong()
        check(lib.otn_query(self.h, C.byref(n), C.byref(p), C.byref(D), C.byref(mm), C.byref(ws)))
        self.n, self.p, self.D, self.m, self.workspace_bytes = n.value, p.value, D.value, mm.value, ws.value
        self.max_batch = max_batch

    def close(self):
        if self.h:
            self._lib.otn_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class StiefelFit:
    """Cost `|| model(xs) - target ||_F` on a product of Stiefel manifolds, evaluated on the GPU.

    model='local': `model_stiefel_local(xs, N, d)` (optimization.py:106-137), isometries (d^2 K, d^2)
    model='full' : `model_Ys_stiefel(xs)`          (optimization.py:147-151), isometries (d^N K, d^N)
    `target` is one (D, D) array (real or complex) or a stack (n_targets, D, D); with several targets each instance of a
    batch selects its own through `target_index`.
    """

    def __init__(self, target, model="local", N=None, d=2, metric="canonical", max_batch=1, device=0, dense=None, factored=False):
        if model not in ("local", "full"):
            raise ValueError("model must be 'local' or 'full'")
        t = np.asarray(target)
        if t.ndim == 2:
            t = t[None]
        if t.ndim != 3 or t.shape[1] != t.shape[2]:
            raise ValueError("target must be (D, D) or (n_targets, D, D)")
        D = t.shape[1]
        if N is None:
            N = int(round(np.log(D) / (2 * np.log(d))))
        if d ** (2 * N) != D:
            raise ValueError(f"target side {D} != d^(2N) = {d ** (2 * N)}")
        self.N, self.d, self.D = N, d, D
        self.model = model
        self.metric = metric if isinstance(metric, str) else metric_alphas(metric)
        metric_id(metric)
        self.device = device
        # dense=True: force the dense D x D chain; False: force the block-diagonal evaluation; None: let the library choose
        # (blocks when the batch is large enough to fill the GPU)
        self.dense = dense
        # factored=True (2-site model, d = 2, N = 4 or 6): never form a D x D layer or chain product; every 16 x 16 local
        # superoperator is applied to the columns of the running product one factor at a time (csrc/factored.cuh)
        self.factored = bool(factored)
        self.max_batch = max_batch
        self._t_re = np.ascontiguousarray(t.real, dtype=np.float64)
        self._t_im = np.ascontiguousarray(t.imag, dtype=np.float64) if np.iscomplexobj(t) and np.any(t.imag) else None
        self.target = np.asarray(target)
        self._handles = {}
        self._tindex = None
        self._tindex_ver = 0
        self._cached = None
        self._row_shard = None
        self.p = d ** N if model == "full" else d * d
        self.last_report = None

    # ------------------------------------------------------------------ plumbing
    def _handle(self, m, K, batch, cplx=False) -> _Handle:
        key = (m, K, bool(cplx))
        h = self._handles.get(key)
        if h is None or h.max_batch < batch:
            if h is not None:
                h.close()
            cap = max(batch, self.max_batch)
            mid = metric_id(self.metric)
            h = _Handle(self.N, self.d, m, K, MODEL_LOCAL if self.model == "local" else MODEL_FULL, max(mid, 0),
                        cap, self._t_re, self._t_im, self.device,
                        _lib.FLAG_COMPLEX if cplx else
                        (_lib.FLAG_FACTORED if self.factored else
                         (0 if self.dense is None else (_lib.FLAG_DENSE if self.dense else _lib.FLAG_BLOCKS))))
            if mid < 0:
                check(h._lib.otn_set_metric_general(h.h, *metric_alphas(self.metric)))
            h.tindex_ver = 0
            h.tindex_len = cap
            self._handles[key] = h
            if self._row_shard is not None:
                self._attach_rows(h)
        return h

    # ------------------------------------------------------------------ row sharding (BASELINE config 5)
    def row_shard(self, group=None):
        """Shard every chain product of this fit by rows over the ranks of `group` (default: the world group of
        `torch.distributed`; one process per GPU of ONE node, `device` = this rank's GPU).  From here on ALL ranks must make the
        same calls with the same arguments (cost / cost_grad / triple / hvp / tr_run ...): the D x D products -- what
        `jax.grad` / `jax.jvp` spend their time on at D = 4096 (stiefel.py:423, 520) -- are split, each tile is stored into all
        peers over NVLink inside the GEMM kernel, the small kernels run replicated.  Results are bitwise equal on every rank
        and to the un-sharded fit.  Needs `dense=True` and D % (64 * world) == 0."""
        import torch.distributed as dist
        if self.dense is not True or self.factored:
            raise ValueError("row_shard needs StiefelFit(..., dense=True): it shards the dense chain products")
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("row_shard needs an initialised torch.distributed process group")
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        if world > 8 or self.D % (64 * world) != 0:
            raise ValueError(f"row_shard: D = {self.D} must be a multiple of 64 * world ({world} ranks, at most 8)")
        self._row_shard = (rank, world, group)
        for h in self._handles.values():
            self._attach_rows(h)
        return self

    def _attach_rows(self, h: "_Handle"):
        import torch.distributed as dist
        rank, world, group = self._row_shard
        if world == 1 or getattr(h, "row_sharded", False):
            return
        mine = C.create_string_buffer(10 * 64)
        check(h._lib.otn_rowshard_export(h.h, mine))
        allh = [None] * world
        dist.all_gather_object(allh, bytes(mine.raw), group=group)
        check(h._lib.otn_rowshard_attach(h.h, rank, world, b"".join(allh)))
        h.row_sharded = True
        dist.barrier(group=group)          # every rank has mapped its peers before the first product stores into them

    def row_shard_ok(self) -> bool:
        """False when a device-side barrier of this rank ever timed out (a peer stopped making the common calls)."""
        ok = True
        for h in self._handles.values():
            if getattr(h, "row_sharded", False):
                v = C.c_int()
                check(h._lib.otn_rowshard_status(h.h, C.byref(v)))
                ok = ok and v.value == 0
        return ok

    def set_target_index(self, index):
        """Target id of each instance of subsequent batched calls (default all 0)."""
        self._tindex = None if index is None else np.ascontiguousarray(index, dtype=np.int32)
        self._tindex_ver += 1

    def _prep(self, xs):
        """-> (array-like [B, m, n, p] float64 contiguous, B, m, K, single, is_cuda)"""
        if _is_cuda_tensor(xs):
            import torch
            if xs.dtype not in (torch.float64, torch.complex128):
                raise TypeError("device tensors must be float64 (or complex128 for complex isometries)")
            if xs.dtype == torch.complex128 and self.model != "full":
                raise NotImplementedError("complex isometries are supported for the full-sites model only (model='full')")
            x = xs.contiguous()
            single = x.dim() == 3
            if single:
                x = x.unsqueeze(0)
            if x.dim() != 4:
                raise ValueError("x must be [B, m, n, p]")
            shape = tuple(x.shape)
            cuda = True
        else:
            if isinstance(xs, (list, tuple)):
                xs = np.stack([np.asarray(a) for a in xs])
            x = np.asarray(xs)
            cplx = False
            if np.iscomplexobj(x):
                if np.abs(x.imag).max() > 0:
                    # complex isometries (the extension the reference marks "TODO: add back the conj", stiefel.py:29-44): full-sites
                    # model through the complex handle (OTN_FLAG_COMPLEX); arrays stay complex128
                    if self.model != "full":
                        raise NotImplementedError("complex isometries are supported for the full-sites model only (model='full'); "
                                                  "the 2-site model's Kronecker embedding is real on the accelerated path")
                    cplx = True
                else:
                    x = x.real
            single = x.ndim == 3
            if single:
                x = x[None]
            if x.ndim != 4:
                raise ValueError("x must be a list of m (n,p) isometries, or an array [m,n,p] / [B,m,n,p]")
            x = np.ascontiguousarray(x, dtype=np.complex128) if cplx else as_f64(x)
            shape = x.shape
            cuda = False
        B, m, n, p = shape
        if p != self.p or n % p != 0:
            raise ValueError(f"isometry shape ({n},{p}) does not fit model '{self.model}' with p = {self.p}")
        return x, B, m, n // p, single, cuda

    def _bind(self, h: _Handle, B):
        """Make the handle's per-instance target ids those of `set_target_index` for the first B instances.  The length that
        was bound is remembered on the handle: a later, larger batch re-validates instead of silently running instances
        beyond it on stale ids."""
        if self._tindex is not None and len(self._tindex) < B:
            raise ValueError(f"target_index has {len(self._tindex)} entries but the batch has {B} instances")
        if h.tindex_ver == self._tindex_ver and B <= h.tindex_len:
            return
        idx = np.zeros(h.max_batch, dtype=np.int32) if self._tindex is None else self._tindex
        nbind = min(len(idx), h.max_batch)
        check(h._lib.otn_set_target_index(h.h, ptr(idx), nbind))
        h.tindex_ver = self._tindex_ver
        h.tindex_len = h.max_batch if self._tindex is None else nbind

    @staticmethod
    def _check_out(arr, shape, cuda, name, cplx=False):
        """User-supplied output arrays go straight to the C ABI as raw addresses: refuse anything that is not a C-contiguous
        float64 array of exactly the expected shape and residency."""
        if _is_cuda_tensor(arr) != cuda:
            raise ValueError(f"{name}: outputs must live where the inputs live ({'device' if cuda else 'host'})")
        if tuple(arr.shape) != tuple(shape):
            raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(arr.shape)}")
        if cuda or hasattr(arr, "is_contiguous"):
            import torch
            if arr.dtype != (torch.complex128 if cplx else torch.float64) or not arr.is_contiguous():
                raise TypeError(f"{name}: need a contiguous {'complex128' if cplx else 'float64'} tensor")
        else:
            if (not isinstance(arr, np.ndarray) or arr.dtype != (np.complex128 if cplx else np.float64) or not arr.flags.c_contiguous
                    or not arr.flags.writeable):
                raise TypeError(f"{name}: need a writeable C-contiguous {'complex128' if cplx else 'float64'} numpy array")

    @staticmethod
    def _empty_like(x, shape, cuda, dtype=np.float64):
        """Uninitialised array living where `x` lives; dtype: np.float64 / np.int32, or "like" = the dtype of x (float64 or
        complex128: gradients of complex isometries are complex)."""
        if cuda:
            import torch
            tdt = x.dtype if dtype == "like" else (torch.float64 if dtype == np.float64 else torch.int32)
            return torch.empty(shape, dtype=tdt, device=x.device)
        return np.empty(shape, dtype=x.dtype if dtype == "like" else dtype)

    def set_metric(self, metric):
        """`metric`: 'euclidean', 'canonical' or a pair (alpha0, alpha1) (stiefel.py:411-426)."""
        a0, a1 = metric_alphas(metric)
        self._cached = None
        self.metric = metric if isinstance(metric, str) else (a0, a1)
        for h in self._handles.values():
            check(h._lib.otn_set_metric_general(h.h, a0, a1))

    def closures(self, metric=None):
        """`(f, retract, gradfunc, hessfunc)` with the signatures `riemannian_trust_region_optimize` expects -- the four
        lambdas a reference notebook builds around its cost closure (experiments/trust_region_pauli_noise.ipynb cell 21).
        Each one is usable on its own and runs on the GPU; they also carry the fit they belong to, so that
        `opentn_b200.trust_region_rcopt.riemannian_trust_region_optimize` recognises the quadruple and runs every
        trust-region step on the device (`otn_tr_run`) instead of driving the four callables from the host."""
        from . import stiefel as _S
        metric = self.metric if metric is None else metric
        metric_id(metric)

        def f(xi):
            return self(xi)

        def retract(xi, eta):
            return _S.retract_stiefel(xi, eta)

        def gradfunc(xi):
            return _S.gradient_stiefel_vec(xi, self, metric=metric)

        def hessfunc(xi):
            return _S.riemannian_hessian_vec(xi, self, metric=metric)

        for fn in (f, retract, gradfunc, hessfunc):
            fn._otn_fit = (self, metric)
        return f, retract, gradfunc, hessfunc

    # ------------------------------------------------------------------ evaluations
    def __call__(self, xs):
        """f(xs) for one instance -> python float (drop-in for the reference's cost closure); for a batch -> array."""
        f = self.cost(xs)
        return f

    def model_matrix(self, xs):
        x, B, m, K, single, cuda = self._prep(xs)
        h = self._handle(m, K, B, _is_cplx(x))
        M = self._empty_like(x, (B, self.D, self.D), cuda)
        self._cached = None
        check(h._lib.otn_model(h.h, ptr(x), B, ptr(M)))
        return M[0] if single else M

    def cost(self, xs):
        x, B, m, K, single, cuda = self._prep(xs)
        h = self._handle(m, K, B, _is_cplx(x))
        self._bind(h, B)
        f = self._empty_like(x, (B,), cuda)
        self._cached = None
        check(h._lib.otn_cost(h.h, ptr(x), B, ptr(f)))
        return float(f[0]) if single else f

    def cost_grad(self, xs, want_egrad=False):
        """-> (f, rgrad[, egrad]); gradients as (n,p) tangent matrices stacked like the input."""
        x, B, m, K, single, cuda = self._prep(xs)
        h = self._handle(m, K, B, _is_cplx(x))
        self._bind(h, B)
        f = self._empty_like(x, (B,), cuda)
        g = self._empty_like(x, tuple(x.shape), cuda, "like")
        e = self._empty_like(x, tuple(x.shape), cuda, "like") if want_egrad else None
        check(h._lib.otn_cost_grad(h.h, ptr(x), B, ptr(f), ptr(g), ptr(e)))
        self._cached = (h, B, object(), _is_cplx(x))
        if single:
            return (float(f[0]), g[0], e[0]) if want_egrad else (float(f[0]), g[0])
        return (f, g, e) if want_egrad else (f, g)

    def triple(self, xs, etas, out=None):
        """Unit of work U1: (f, rgrad, H eta) sharing the primal pass."""
        x, B, m, K, single, cuda = self._prep(xs)
        e, B2, m2, K2, _, cuda2 = self._prep(etas)
        if (B, m, K, cuda) != (B2, m2, K2, cuda2):
            raise ValueError("xs and etas must have the same shape and residency")
        x, e = _promote_pair(x, e)
        h = self._handle(m, K, B, _is_cplx(x))
        self._bind(h, B)
        if out is None:
            f = self._empty_like(x, (B,), cuda)
            g = self._empty_like(x, tuple(x.shape), cuda, "like")
            hv = self._empty_like(x, tuple(x.shape), cuda, "like")
        else:
            f, g, hv = out
            self._check_out(f, (B,), cuda, "out[0] (f)")
            self._check_out(g, tuple(x.shape), cuda, "out[1] (rgrad)", _is_cplx(x))
            self._check_out(hv, tuple(x.shape), cuda, "out[2] (heta)", _is_cplx(x))
        check(h._lib.otn_triple(h.h, ptr(x), ptr(e), B, ptr(f), ptr(g), ptr(hv)))
        self._cached = (h, B, object(), _is_cplx(x))
        if single:
            return float(f[0]), g[0], hv[0]
        return f, g, hv

    def triple_submit(self, xs, etas, out):
        """Streamed form of `triple` for host arrays [B,m,n,p] (pinned torch tensors or numpy): enqueues the copies and kernels
        of this batch and returns a ticket at once; `out = (f, rgrad, heta)` are host arrays that `triple_wait(ticket)` fills.
        Submitting the next batch before waiting overlaps its copies with this batch's kernels (otn_triple_submit).
        With CUDA tensors throughout (inputs and outputs) nothing is copied: the submission only skips the host synchronisation
        of `triple`, so consecutive batches overlap across their boundary; results are valid after `triple_wait`."""
        x, B, m, K, single, cuda = self._prep(xs)
        e, B2, m2, K2, _, cuda2 = self._prep(etas)
        if (B, m, K) != (B2, m2, K2):
            raise ValueError("xs and etas must have the same shape")
        x, e = _promote_pair(x, e)
        outs_cuda = [bool(getattr(o, "is_cuda", False)) for o in out]
        if len({cuda, cuda2, *outs_cuda}) != 1:
            raise ValueError("triple_submit: the arrays of one submission must be all host or all device resident")
        h = self._handle(m, K, B, _is_cplx(x))
        self._bind(h, B)
        f, g, hv = out
        self._check_out(f, (B,), cuda, "out[0] (f)")
        self._check_out(g, tuple(x.shape), cuda, "out[1] (rgrad)", _is_cplx(x))
        self._check_out(hv, tuple(x.shape), cuda, "out[2] (heta)", _is_cplx(x))
        ticket = C.c_longlong(-1)
        check(h._lib.otn_triple_submit(h.h, ptr(x), ptr(e), B, ptr(f), ptr(g), ptr(hv), C.byref(ticket)))
        self._cached = None
        self._inflight = getattr(self, "_inflight", {})
        key = (id(h), ticket.value)                        # tickets are per handle (one handle per (m, K, max_batch))
        self._inflight[key] = (h, x, e, out)               # keep the host arrays alive until the wait
        return key

    def triple_wait(self, ticket=None):
        """Block until the outputs of `ticket` (default: every outstanding submission) are in their host arrays."""
        inflight = getattr(self, "_inflight", {})
        if ticket is None:
            for t in sorted(inflight):
                self.triple_wait(t)
            return
        if ticket not in inflight:
            raise OtnError(f"ticket {ticket} is not outstanding")
        h = inflight[ticket][0]
        check(h._lib.otn_triple_wait(h.h, ticket[1]))
        del inflight[ticket]

    def hvp_cached(self, etas):
        """H eta at the x of the latest cost_grad / triple call."""
        cached = getattr(self, "_cached", None)
        if cached is None:
            raise OtnError("no primal cache: call cost_grad or triple first")
        h, B = cached[0], cached[1]
        e, B2, m2, K2, single, cuda = self._prep(etas)
        if B2 != B:
            raise ValueError("batch differs from the cached primal pass")
        if cached[3] and not _is_cplx(e):                       # primal pass was complex: a real direction is promoted
            if isinstance(e, np.ndarray):
                e = np.ascontiguousarray(e, dtype=np.complex128)
            else:
                import torch
                e = e.to(torch.complex128).contiguous()
        elif _is_cplx(e) and not cached[3]:
            raise ValueError("complex direction for a real cached primal pass")
        hv = self._empty_like(e, tuple(e.shape), cuda, "like")
        check(h._lib.otn_hvp_cached(h.h, ptr(e), B, ptr(hv)))
        return hv[0] if single else hv

    def hvp(self, xs, etas):
        return self.triple(xs, etas)[2]

    # ------------------------------------------------------------------ batched trust region
    def tr_run(self, xs, niter=20, rho_trust=0.125, radius_init=0.01, maxradius=0.1, tcg_maxiter=0, tcg_abstol=1e-8,
               tcg_reltol=1e-6, save_x=False, radius=None):
        """Batched device-side trust region.  Returns (x_final, report dict).  `xs` [B,m,n,p] (or one instance)."""
        x, B, m, K, single, cuda = self._prep(xs)
        if _is_cplx(x):
            raise NotImplementedError("the batched device solver is real-only; for complex isometries run "
                                      "opentn_b200.trust_region_rcopt.riemannian_trust_region_optimize over "
                                      "opentn_b200.complex_stiefel.closures(fit)")
        h = self._handle(m, K, B, _is_cplx(x))
        self._bind(h, B)
        if cuda:
            xw = x.clone()
        else:
            xw = x.copy()
        prm = _lib.TrParams(rho_trust, radius_init, maxradius, niter, tcg_maxiter, tcg_abstol, tcg_reltol)
        f_iter = np.empty((niter, B)); rho = np.empty((niter, B)); rad = np.empty((niter, B))
        n_cg = np.empty((niter, B), dtype=np.int32); onb = np.empty((niter, B), dtype=np.int32); acc = np.empty((niter, B), dtype=np.int32)
        status = np.zeros(B, dtype=np.int32)
        x_iter = np.empty((niter + 1,) + tuple(x.shape)) if save_x else None
        rad_io = None
        if radius is not None:
            rad_io = np.ascontiguousarray(np.broadcast_to(np.asarray(radius, dtype=np.float64), (B,)).copy())
        else:
            rad_io = np.full(B, radius_init)
        self._cached = None
        rep = _lib.TrReport(ptr(f_iter), ptr(rho), ptr(rad), ptr(n_cg), ptr(onb), ptr(acc), ptr(x_iter), ptr(rad_io), ptr(status), 0)
        check(h._lib.otn_tr_run(h.h, ptr(xw), B, C.byref(prm), C.byref(rep)))
        report = dict(f_iter=f_iter, rho=rho, radius_iter=rad, n_cg=n_cg, on_boundary=onb, accepted=acc, status=status,
                      radius=rad_io, hvp_total=rep.hvp_total, x_iter=x_iter)
        self.last_report = report
        return (xw[0] if single else xw), report

    def close(self):
        sharded = [h for h in self._handles.values() if getattr(h, "row_sharded", False)]
        if sharded and self._row_shard is not None:
            # collective: every rank unmaps its peers' arrays before any rank frees its own
            import torch.distributed as dist
            group = self._row_shard[2]
            dist.barrier(group=group)
            for h in sharded:
                check(h._lib.otn_rowshard_detach(h.h))
                h.row_sharded = False
            dist.barrier(group=group)
        for h in self._handles.values():
            h.close()
        self._handles = {}
