"""BASELINE config 4: tau x rank x seed sweep of 2-site Kitaev fits (N=4, m=9), sharded over the ranks of a torchrun launch.
    python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/c4_sweep.py [seeds_per_cell] [niter]
"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
from opentn_b200 import sharding, sweep
from opentn_b200 import transformations as T
T.set_device(lr)

seeds_per_cell = int(sys.argv[1]) if len(sys.argv) > 1 else 256
niter = int(sys.argv[2]) if len(sys.argv) > 2 else 5
taus, ranks = [0.25, 0.5, 1.0, 2.0], [2, 4, 8, 16]
Lnn = [T.get_kitaev_nn_linbladian(1.0)]
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
sweep.run_sweep(Lnn, 4, 2, taus, ranks, [0, 1], n_steps=4, niter=1, device=lr, rank=rank, world=world, gather=world > 1)   # warm-up
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
tm = {}
res = sweep.run_sweep(Lnn, 4, 2, taus, ranks, list(range(seeds_per_cell)), n_steps=4, niter=niter, device=lr, rank=rank, world=world,
                      gather=world > 1, timings=tm)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
wall = time.perf_counter() - t0
if rank == 0:
    n_inst = sum(len(r["f0"]) for r in res.values())
    print(f"world={world}: {n_inst} instances x {niter} TR iterations in {wall:.3f} s = {n_inst * niter / wall:.0f} TR iterations/s "
          f"(rank 0: set-up {tm['setup_s']:.3f} s, solve {tm['solve_s']:.3f} s)")
    for K, r in res.items():
        for ti, tau in enumerate(taus):
            sel = r["tau_index"] == ti
            print(f"  K={K:2d} tau={tau:4.2f}: f0 median {np.median(r['f0'][sel]):.3e} -> f_final median {np.median(r['f_final'][sel]):.3e}"
                  f"  (improved {np.mean(r['f_final'][sel] < r['f0'][sel]) * 100:.0f}%, status!=0: {int((r['status'][sel] != 0).sum())})")
if world > 1:
    dist.destroy_process_group()
