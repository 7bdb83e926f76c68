"""Per-category timing of the 2-site (local) model: C2 (B=1) and C4-like batches (developer tool)."""
import ctypes as C
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from opentn_b200 import StiefelFit, _lib
from opentn_b200 import transformations as T
from opentn_b200 import optimization as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
K = int(sys.argv[2]) if len(sys.argv) > 2 else 16
m = int(sys.argv[3]) if len(sys.argv) > 3 else 9
MODE = sys.argv[4] if len(sys.argv) > 4 else "auto"          # auto | dense | blocks | factored
N, d = 4, 2
target = T.exp_operator_dt(T.create_kitaev_liouvillians(N, d, 1.0, pbc=True)[0], 0.5)
xs0 = np.array([T.super2ortho(x.real, rank=K) for x in O.get_kitaev_trotter_local_ansatz(1.0, 0.5, (m - 1) // 2)])
rng = np.random.default_rng(0)
xs = np.repeat(xs0[None], B, axis=0)
z = rng.standard_normal(xs.shape)
xtz = np.swapaxes(xs, -1, -2) @ z
et = z - xs @ (0.5 * (xtz + np.swapaxes(xtz, -1, -2)))
x = torch.from_numpy(xs).cuda(); e = torch.from_numpy(et).cuda()
kw = {"auto": {}, "dense": dict(dense=True), "blocks": dict(dense=False), "factored": dict(factored=True)}[MODE]
fit = StiefelFit(target, model="local", N=N, d=d, max_batch=B, **kw)
h = fit._handle(m, K, B)
lib = _lib.load()
names = ["gemm", "assembly", "back_assembly", "riem", "tr", "other"]
for what in ("triple", "tr"):
    if what == "triple":
        fn = lambda: fit.triple(x, e)
    else:
        fn = lambda: fit.tr_run(x, niter=2)
    fn(); torch.cuda.synchronize()
    lib.otn_profile_enable(h.h, 1)
    t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); wall = (time.perf_counter() - t0) * 1e3
    ms = (C.c_double * 6)(); n = (C.c_longlong * 6)(); fl = (C.c_double * 6)()
    lib.otn_profile_read(h.h, ms, n, fl); lib.otn_profile_enable(h.h, 0)
    print(f"--- {what} [{MODE}]: B={B} K={K} m={m}: wall {wall:.2f} ms", ("n_cg mean %.2f" % r[1]["n_cg"].mean()) if what == "tr" else "")
    for k in range(6):
        if n[k]:
            print(f"{names[k]:14s} {ms[k]:8.2f} ms  {n[k]:5d} launches" + (f"  {fl[k]/ms[k]/1e9:.1f} TFLOP/s" if k == 0 else ""))
