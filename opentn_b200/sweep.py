"""tau x Kraus-rank x seed sweeps of 2-site trust-region fits (BASELINE config 4), sharded by instance.

An instance = (model Lindbladian, tau, K, seed).  Instances are grouped by K (a batch must share the isometry shape), every
group is one `StiefelFit` with one target per tau and a per-instance target index, each rank runs `b mod world == rank`
of every group on its GPU with the batched device solver, and the only collective is the final all-gather of costs, status
flags and isometries (`opentn_b200.sharding`)."""
from __future__ import annotations

import numpy as np

from . import transformations as T
from .fit import StiefelFit
from .optimization import get_general_trotter_local_ansatz
from .stiefel import polar_decomposition_rectangular, project


def kicked_start(xs0, seed, kick=1e-2):
    """Trotter starting point retracted after a random tangent kick of canonical norm `kick` (default_rng(seed))."""
    rng = np.random.default_rng(seed)
    eta = np.stack([rng.standard_normal(x.shape) for x in xs0])
    xs0 = np.stack(xs0)
    eta = project(xs0, eta)
    xte = np.swapaxes(xs0, -1, -2) @ eta
    nrm = np.sqrt(np.sum(eta * eta) - 0.5 * np.sum(xte * xte))
    return polar_decomposition_rectangular(xs0, eta * (kick / nrm))


def kicked_starts(xs0_stack, seeds, kick=1e-2):
    """`kicked_start` for many instances at once: xs0_stack (B, m, n, p), one seed per instance.  The random directions are drawn
    per instance exactly as in `kicked_start`; projection and polar retraction are one batched device call each."""
    xs0_stack = np.asarray(xs0_stack, dtype=np.float64)
    eta = np.empty_like(xs0_stack)
    for b, seed in enumerate(seeds):
        rng = np.random.default_rng(seed)
        for l in range(xs0_stack.shape[1]):
            eta[b, l] = rng.standard_normal(xs0_stack.shape[2:])
    eta = project(xs0_stack, eta)
    xte = np.swapaxes(xs0_stack, -1, -2) @ eta
    nrm = np.sqrt(np.sum(eta * eta, axis=(1, 2, 3)) - 0.5 * np.sum(xte * xte, axis=(1, 2, 3)))
    return polar_decomposition_rectangular(xs0_stack, eta * (kick / nrm)[:, None, None, None])


def _seed_array(seeds):
    return np.ascontiguousarray(np.asarray(seeds, dtype=np.uint64))


def kicked_starts_device(xs0_stack, seeds, kick=1e-2, device=None):
    """The same construction entirely on the device (otn_kicked_starts): the directions come from the counter-based generator of
    the library (Philox4x32-10, one independent stream per seed; convention in include/otn.h) instead of numpy's PCG64, so nothing
    is drawn in a host loop.  xs0_stack: numpy array or CUDA tensor (B, m, n, p); returns the same kind."""
    from . import _lib
    from ._lib import check, ptr
    seeds = _seed_array(seeds)
    cuda = hasattr(xs0_stack, "is_cuda") and xs0_stack.is_cuda
    if cuda:
        x0 = xs0_stack.contiguous()
        import torch
        out = torch.empty_like(x0)
        dev = x0.device.index if device is None else device
    else:
        x0 = np.ascontiguousarray(xs0_stack, dtype=np.float64)
        out = np.empty_like(x0)
        dev = T._DEVICE if device is None else device
    B, m, n, p = x0.shape
    assert len(seeds) == B, "one seed per instance"
    check(_lib.load().otn_kicked_starts(ptr(x0), seeds.ctypes.data, B, m, n, p, float(kick), ptr(out), dev))
    return out


def random_stiefel_device(seeds, m, n, p, device=None):
    """Haar-random isometries (B, m, n, p) drawn on the device (otn_random_stiefel); numpy array out."""
    from . import _lib
    from ._lib import check, ptr
    seeds = _seed_array(seeds)
    out = np.empty((len(seeds), m, n, p))
    check(_lib.load().otn_random_stiefel(seeds.ctypes.data, len(seeds), m, n, p, ptr(out), T._DEVICE if device is None else device))
    return out


def plan_cells(ranks, n_taus, world, rank):
    """(K, tau index) cells of the sweep owned by `rank`.  The K-major cell list is dealt out in a snake (0, 1, .., w-1, w-1, .., 1,
    0, 0, 1, ..): every rank gets the same number of cells, and cheap cells (small K, i.e. small isometries) are paired with
    expensive ones, so the ranks finish together.  Within a rank, cells of the same K form one batch (one handle per shape)."""
    cells = [(K, ti) for K in ranks for ti in range(n_taus)]
    mine = []
    for i, cell in enumerate(cells):
        j = i % (2 * world)
        if (j if j < world else 2 * world - 1 - j) == rank:
            mine.append(cell)
    return mine


def make_fit(lindbladians, N, d, taus, device=0, pbc=True):
    """The cost object of a sweep: targets exp(tau L) for every tau (device Liouvillian assembly + one batched expm), one handle per
    isometry shape created on first use.  Reuse it across `run_sweep` calls to keep the handles (work buffers) alive."""
    T.set_device(device)
    Lvec = T.create_2local_liouvillians(lindbladians, N, d, pbc=pbc, backend="cuda")[0]
    return StiefelFit(T.exp_operator_sweep(Lvec, taus), model="local", N=N, d=d, metric="canonical", max_batch=1, device=device)


def instance_seed(K, tau_index, seed):
    """64-bit generator seed of instance (K, tau, seed): independent of how the sweep is sharded."""
    return (int(K) << 48) | (int(tau_index) << 32) | (int(seed) & 0xFFFFFFFF)


def run_sweep(lindbladians, N, d, taus, ranks, seeds, n_steps=4, niter=10, device=0, rank=0, world=1, pbc=True, max_batch=4096,
              gather=False, fit=None, timings=None, kick=1e-2, to_host=True,
              gather_names=("tau_index", "seed", "f0", "f_final", "status", "x_final")):
    """tau x rank x seed sweep of 2-site trust-region fits (BASELINE config 4).  Returns {K: dict(tau_index, seed, f0, f_final,
    x0, x_final, status)} for the instances of the cells owned by `rank` (`plan_cells`); with `gather=True` (inside an initialised
    torch.distributed job) the final all-gather follows and every rank returns all instances of every K.

    Everything between the Lindbladian and the final costs runs on the device: Liouvillian assembly (otn_liouvillian), one batched
    expm over the tau grid (otn_expm_batch), Choi factorisation of the Trotter layers (otn_super2ortho), kicked starts from the
    counter-based generator (otn_kicked_starts), the batched trust region (otn_tr_run).  `fit`: a StiefelFit to reuse (its targets
    must be exp(tau L) for `taus`); `timings`: dict that receives 'setup_s' and 'solve_s' (host wall clock around synchronised
    device work; 'gather_s' with `gather`).  `to_host=False` returns CUDA tensors instead of numpy arrays; `gather_names`: the arrays
    the final all-gather carries (costs, status and final isometries by default; the starts `x0` stay on the rank that drew them)."""
    import time

    import torch
    T.set_device(device)
    dev = torch.device("cuda", device)
    t0 = time.perf_counter()
    own_fit = fit is None
    if own_fit:
        fit = make_fit(lindbladians, N, d, taus, device=device, pbc=pbc)
    cells = plan_cells(ranks, len(taus), world, rank)
    # Trotter layers of every tau this rank needs (16 x 16 expm on the host: microseconds), factorised per K on the device
    need_taus = sorted({ti for _, ti in cells})
    layers = {ti: np.stack([np.asarray(x).real for x in get_general_trotter_local_ansatz(lindbladians, taus[ti], n_steps)]) for ti in need_taus}
    groups = {}
    for K, ti in cells:
        groups.setdefault(K, []).append(ti)
    seeds = list(seeds)
    prepared = {}
    for K, tis in groups.items():
        m_layers = layers[tis[0]].shape[0]
        flat = T.super2ortho_batch(np.concatenate([layers[ti] for ti in tis]), K)            # (len(tis) * m, 4K, 4)
        starts = torch.from_numpy(flat.reshape(len(tis), m_layers, 4 * K, d * d)).to(dev)
        tindex = np.repeat(np.asarray(tis, dtype=np.int32), len(seeds))
        sd = np.array([instance_seed(K, ti, s) for ti in tis for s in seeds], dtype=np.uint64)
        x0 = starts[torch.arange(len(tis), device=dev).repeat_interleave(len(seeds))].contiguous()
        prepared[K] = (kicked_starts_device(x0, sd, kick), tindex, np.tile(np.asarray(seeds), len(tis)))
    torch.cuda.synchronize(dev)
    t1 = time.perf_counter()
    out = {}
    for K, (xs, tindex, seed_arr) in prepared.items():
        B = xs.shape[0]
        xf = torch.empty_like(xs)
        f0 = np.empty(B); ff = np.empty(B); status = np.zeros(B, dtype=np.int32)
        for c0 in range(0, B, max_batch):
            sl = slice(c0, min(c0 + max_batch, B))
            fit.set_target_index(tindex[sl])
            xe, rep = fit.tr_run(xs[sl], niter=niter)
            ff[sl] = fit.cost(xe).cpu().numpy()
            f0[sl] = rep["f_iter"][0]
            status[sl] = rep["status"]
            xf[sl] = xe
        out[K] = dict(tau_index=tindex, seed=seed_arr, f0=f0, f_final=ff, x0=xs, x_final=xf, status=status)
    torch.cuda.synchronize(dev)
    t2 = time.perf_counter()
    if timings is not None:
        timings["setup_s"], timings["solve_s"] = t1 - t0, t2 - t1
    if own_fit:
        fit.close()
    res = {}
    for K, r in out.items():
        r = {k: (v if torch.is_tensor(v) else torch.from_numpy(np.ascontiguousarray(v)).to(dev)) for k, v in r.items()}
        res[K] = r
    if gather:
        t3 = time.perf_counter()
        res = gather_groups(res, list(ranks), len(taus), len(seeds), world, names=gather_names)
        torch.cuda.synchronize(dev)
        if timings is not None:
            timings["gather_s"] = time.perf_counter() - t3
    if not to_host:
        return res
    return {K: {k: v.cpu().numpy() for k, v in r.items()} for K, r in res.items()}


def gather_groups(res, ranks, n_taus, n_seeds, world, names=None):
    """Final all-gather of a cell-partitioned sweep (the one collective of the path): every rank contributes the K groups it owns;
    on return every rank holds all instances of every K, ordered by (tau index, seed).  `res`: {K: {name: CUDA tensor}}."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or world == 1:
        return res
    some = next(iter(res.values()))
    dev = some["f0"].device
    # Layout of every rank's contribution, known to all ranks from the cell plan: K groups in the order of `ranks`, each
    # counts[r][K] instances of width_K float64 values (all arrays of the group packed instance by instance; integers are exactly
    # representable: indices, 32-bit seeds, status bits).  ONE collective carries the whole sweep.
    plans = [plan_cells(ranks, n_taus, world, r) for r in range(world)]
    counts = {K: [sum(1 for (kk, _) in plans[r] if kk == K) * n_seeds for r in range(world)] for K in ranks}
    fields, width = {}, {}
    for K in ranks:
        mine = res.get(K)
        fields[K] = [(name, tuple(mine[name].shape[1:]) if mine is not None else tuple(_tail_shape(name, like, K, some)), like.dtype)
                     for name, like in some.items() if names is None or name in names]
        width[K] = [int(np.prod(shape)) if shape else 1 for _, shape, _ in fields[K]]
    offs = [{} for _ in range(world)]
    for r in range(world):
        o = 0
        for K in ranks:
            offs[r][K] = o
            o += counts[K][r] * sum(width[K])
        offs[r][None] = o
    cap = max(offs[r][None] for r in range(world))
    me = dist.get_rank()
    flat = torch.zeros(max(cap, 1), dtype=torch.float64, device=dev)
    for K in ranks:
        mine = res.get(K)
        if mine is None or counts[K][me] == 0:
            continue
        n = counts[K][me]
        rows = torch.cat([mine[name].reshape(n, w).to(torch.float64) for (name, _, _), w in zip(fields[K], width[K])], dim=1)
        flat[offs[me][K]: offs[me][K] + rows.numel()] = rows.reshape(-1)
    parts = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(parts, flat)                         # (list form: also available on gloo, which the CPU tests use)
    out = {}
    for K in ranks:
        wk = sum(width[K])
        full = torch.cat([parts[r][offs[r][K]: offs[r][K] + counts[K][r] * wk].view(counts[K][r], wk)
                          for r in range(world) if counts[K][r]])
        merged, o = {}, 0
        for (name, shape, dtype), w in zip(fields[K], width[K]):
            merged[name] = full[:, o:o + w].reshape((full.shape[0],) + shape).to(dtype)
            o += w
        if "tau_index" in merged and "seed" in merged:        # canonical order whatever the partition: by (tau index, seed)
            key = merged["tau_index"].to(torch.int64) * (1 << 32) + merged["seed"].to(torch.int64)
            order = torch.argsort(key)
            merged = {name: v[order] for name, v in merged.items()}
        else:
            merged = {name: v.contiguous() for name, v in merged.items()}
        out[K] = merged
    return out


def _tail_shape(name, like, K, some):
    """Trailing shape of array `name` for Kraus rank K given an example of another rank (isometries are (m, 4K, 4))."""
    if like.dim() <= 1:
        return ()
    m, _, p = like.shape[1:]
    return (m, p * K, p)
