"""Per-category CUDA-event timing of the batched trust region on the headline workload (developer tool)."""
import ctypes as C
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from opentn_b200 import StiefelFit, _lib
from opentn_b200 import transformations as T
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
niter = int(sys.argv[2]) if len(sys.argv) > 2 else 3
target = T.exp_operator_dt(T.create_kitaev_liouvillians(4, 2, 1.0, pbc=False)[0], 4.0)
xs, et = bench.make_inputs(0, B)
x = torch.from_numpy(xs).cuda()
fit = StiefelFit(target, model="full", N=4, d=2, max_batch=B)
h = fit._handle(3, 16, B)
lib = _lib.load()
fit.tr_run(x, niter=1)
lib.otn_profile_enable(h.h, 1)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
_, rep = fit.tr_run(x, niter=niter)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) * 1e3
ms = (C.c_double * 6)(); n = (C.c_longlong * 6)(); fl = (C.c_double * 6)()
lib.otn_profile_read(h.h, ms, n, fl)
names = ["gemm", "assembly", "back_assembly", "riem", "tr", "other"]
print(f"wall {wall:.1f} ms for {niter} iterations, n_cg mean {rep['n_cg'].mean():.2f}")
for k in range(6):
    print(f"{names[k]:14s} {ms[k]:8.2f} ms  {n[k]:5d} launches")
print("sum", sum(ms))
