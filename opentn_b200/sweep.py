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


def run_sweep(lindbladians, N, d, taus, ranks, seeds, n_steps=4, niter=10, device=0, rank=0, world=1, pbc=True, max_batch=1024,
              gather=None):
    """Returns {K: dict(tau_index, seed, f0, f_final, x0, x_final, status)} for the instances owned by `rank` (or, with
    `gather` = a function like sharding.gather_instances, for all instances on every rank)."""
    import torch
    # set-up on the device: Liouvillian assembly (otn_liouvillian) and one batched expm over the tau grid (otn_expm_batch)
    T.set_device(device)
    Lvec = T.create_2local_liouvillians(lindbladians, N, d, pbc=pbc, backend="cuda")[0]
    targets = T.exp_operator_sweep(Lvec, taus)
    out = {}
    for K in ranks:
        # starting points: one batched device super2ortho over the Trotter layers of every tau (otn_super2ortho)
        layers = [np.stack([np.asarray(x).real for x in get_general_trotter_local_ansatz(lindbladians, tau, n_steps)]) for tau in taus]
        m_layers = layers[0].shape[0]
        flat = T.super2ortho_batch(np.concatenate(layers), K)
        starts = [flat[i * m_layers:(i + 1) * m_layers] for i in range(len(taus))]
        inst = [(ti, s) for ti in range(len(taus)) for s in seeds]
        mine = inst[rank::world]
        if not mine:
            continue
        xs = kicked_starts(np.stack([starts[ti] for ti, _ in mine]), [s for _, s in mine])
        tindex = np.array([ti for ti, _ in mine], dtype=np.int32)
        xf = np.empty_like(xs)
        f0 = np.empty(len(mine)); ff = np.empty(len(mine)); status = np.zeros(len(mine), dtype=np.int32)
        for c0 in range(0, len(mine), max_batch):
            sl = slice(c0, min(c0 + max_batch, len(mine)))
            fit = StiefelFit(targets, model="local", N=N, d=d, metric="canonical", max_batch=sl.stop - sl.start, device=device)
            fit.set_target_index(tindex[sl])
            xd = torch.from_numpy(xs[sl]).to(torch.device("cuda", device))
            xe, rep = fit.tr_run(xd, niter=niter)
            ff[sl] = fit.cost(xe).cpu().numpy()
            f0[sl] = rep["f_iter"][0]
            status[sl] = rep["status"]
            xf[sl] = xe.cpu().numpy()
            fit.close()
        res = dict(tau_index=tindex, seed=np.array([s for _, s in mine]), f0=f0, f_final=ff, x0=xs, x_final=xf, status=status)
        if gather is not None:
            total = len(inst)
            dev = torch.device("cuda", device)
            res = {k: gather(torch.from_numpy(np.ascontiguousarray(v)).to(dev), total).cpu().numpy() for k, v in res.items()}
        out[K] = res
    return out
