"""Small end-to-end exercise of every kernel family for compute-sanitizer (developer tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from opentn_b200 import StiefelFit
from opentn_b200 import stiefel as S, transformations as T, optimization as O

rng = np.random.default_rng(0)
def rs(n, p):
    q, r = np.linalg.qr(rng.standard_normal((n, p))); return q * np.sign(np.diag(r))
Tt = rng.standard_normal((256, 256)) * 0.1
for model, p, K, m in (("full", 16, 5, 3), ("full", 16, 16, 2), ("local", 4, 3, 3)):
    for dense in (False, True):
        xs = np.array([[rs(p * K, p) for _ in range(m)] for _ in range(2)])
        et = S.project(xs, rng.standard_normal(xs.shape))
        fit = StiefelFit(Tt, model=model, N=4, d=2, max_batch=2, dense=dense)
        f, g, h = fit.triple(xs, et)
        fit.cost(xs); fit.cost_grad(xs); fit.hvp_cached(et); fit.model_matrix(xs)
        x, rep = fit.tr_run(xs, niter=1)
        fit.close()
        print(model, dense, f)
x = rs(40, 4)
T.ortho2super(x); T.local2full(rng.standard_normal((16, 16)), 4, 2, True); T.compose_superops_list([Tt, Tt, Tt])
S.polar_decomposition_rectangular(x, 0.1 * S.project(x, rng.standard_normal(x.shape))); S.canonical_metric(x, x, x)
S.riemannian_connection(x, x, x, x, 1, 0.5); O.frobenius_norm(Tt, Tt + 1j)
T64 = rng.standard_normal((64, 64))
xs = np.array([[rs(8 * 3, 8) for _ in range(2)]])
fit = StiefelFit(T64, model="full", N=3, d=2); fit.triple(xs, S.project(xs, rng.standard_normal(xs.shape))); fit.close()
print("done")
