"""One dense N=6 unit of work (cost + Riemannian gradient + HVP, D = 4096, K = 16, m = 3) on one GPU: the kernels of BASELINE config
5's fit (dense chain GEMM at D = 4096, embedding, three-factor back-embedding) for an ncu capture:
    ncu --set full --clock-control none --import-source on -k 'regex:back_embed3|embed_kernel|gemm_dmma' -c 12 -o out python tools/n6_dense_u1.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opentn_b200 import StiefelFit                      # noqa: E402
from opentn_b200 import transformations as T            # noqa: E402

full = T.create_kitaev_liouvillians(6, 2, 1.0, pbc=True, backend="cuda")[0]
target = np.ascontiguousarray(T.exp_operator_dt(full, 4.0, backend="cuda").real)
rng = np.random.default_rng(6)
xs = np.stack([np.linalg.qr(rng.standard_normal((64, 4)))[0] for _ in range(3)])[None]
et = rng.standard_normal(xs.shape)
et = et - xs @ (0.5 * (np.swapaxes(xs, -1, -2) @ et + np.swapaxes(et, -1, -2) @ xs))
xd, ed = torch.from_numpy(xs).cuda(), torch.from_numpy(et).cuda()
fit = StiefelFit(target, model="local", N=6, d=2, metric="canonical", dense=True)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for _ in range(reps):
    f, g, h = fit.triple(xd, ed)
torch.cuda.synchronize()
print("f =", float(f[0]))
fit.close()
