"""N = 6 single-instance triple (f, rgrad, H eta): block-diagonal chain vs factored evaluation (developer tool)."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from opentn_b200 import StiefelFit
from opentn_b200 import transformations as T
from opentn_b200 import optimization as O
from opentn_b200 import stiefel as S

K = int(sys.argv[1]) if len(sys.argv) > 1 else 16
Lnn = T.get_kitaev_nn_linbladian(1.0)
L6 = T.create_2local_liouvillians(Lnn, 6, 2, pbc=True, backend="cuda")[0]
T6 = T.exp_operator_dt(L6, 4.0, backend="cuda")
xs = np.array([T.super2ortho(x.real, rank=K) for x in O.get_general_trotter_local_ansatz([Lnn], 4, n=1)])
rng = np.random.default_rng(0)
et = np.array([S.project(x, rng.standard_normal(x.shape)) for x in xs])
x = torch.from_numpy(xs[None]).cuda(); e = torch.from_numpy(et[None]).cuda()
for mode, kw in (("blocks", dict(dense=False)), ("factored", dict(factored=True))):
    fit = StiefelFit(T6, model="local", N=6, d=2, **kw)
    for what, fn in (("cost", lambda: fit.cost(x)), ("cost_grad", lambda: fit.cost_grad(x)), ("triple", lambda: fit.triple(x, e)),
                     ("tr_iteration", lambda: fit.tr_run(x, niter=1))):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        print(f"N=6 K={K} {mode:9s} {what:12s} {(time.perf_counter() - t0) / reps * 1e3:9.2f} ms")
    fit.close()

# per-category device time of one factored triple
import ctypes as C
from opentn_b200 import _lib
fit = StiefelFit(T6, model="local", N=6, d=2, factored=True)
fit.triple(x, e); torch.cuda.synchronize()
h = fit._handle(3, K, 1)
lib = _lib.load()
lib.otn_profile_enable(h.h, 1)
fit.triple(x, e); torch.cuda.synchronize()
ms = (C.c_double * 6)(); n = (C.c_longlong * 6)(); fl = (C.c_double * 6)()
lib.otn_profile_read(h.h, ms, n, fl); lib.otn_profile_enable(h.h, 0)
for k, nm in enumerate(["factored kernel", "assembly", "back_assembly", "riem", "tr", "other"]):
    if n[k]:
        print(f"  {nm:16s} {ms[k]:7.3f} ms {n[k]:3d} launches" + (f"  {fl[k] / ms[k] / 1e9:.1f} TFLOP/s" if k == 0 else ""))
