"""Single-instance trust-region latency (BASELINE configs 1 and 2: the reference's own notebook runs), wall clock per TR
iteration with and without the CUDA graph of the tCG step.  Usage: python tools/tr_latency.py"""
import os
import subprocess
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if len(sys.argv) > 1:
    import numpy as np
    import torch
    from opentn_b200 import StiefelFit
    from opentn_b200 import transformations as T
    from opentn_b200 import optimization as O
    for name, Li, tau, n, K in (("C1 pspl m=3 K=10", [np.sqrt(1.0) * (np.kron(T.X, T.I) - np.kron(T.I, T.X)), np.sqrt(1.0) * (np.kron(T.Z, T.I) - np.kron(T.I, T.Z))], 1.0, 1, 10),
                                ("C2 pspl m=9 K=10", [np.sqrt(1.0) * (np.kron(T.X, T.I) - np.kron(T.I, T.X)), np.sqrt(1.0) * (np.kron(T.Z, T.I) - np.kron(T.I, T.Z))], 0.5, 4, 10)):
        target = T.exp_operator_dt(T.create_2local_liouvillians(Li, 4, 2, pbc=True)[0], tau)
        xs0 = np.array([T.super2ortho(x.real, rank=K) for x in O.get_general_trotter_local_ansatz(Li, tau, n)])
        fit = StiefelFit(target, model="local", N=4, d=2)
        x = torch.from_numpy(xs0[None]).cuda()
        fit.tr_run(x, niter=2); torch.cuda.synchronize()
        niter = 10
        t0 = time.perf_counter()
        xf, rep = fit.tr_run(x, niter=niter)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"{sys.argv[1]:9s} {name}: {dt / niter * 1e3:7.2f} ms per TR iteration, n_cg mean {rep['n_cg'].mean():.1f}, "
              f"{dt / (rep['n_cg'].sum() + niter) * 1e6:6.1f} us per tCG step, f {rep['f_iter'][0, 0]:.6f} -> {rep['f_iter'][-1, 0]:.6f}")
        fit.close()
else:
    for tag, env in (("graph", {}), ("no-graph", {"OTN_NO_GRAPH": "1"})):
        subprocess.run([sys.executable, __file__, tag], env={**os.environ, **env}, check=True)
