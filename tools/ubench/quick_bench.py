import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from oracle import port as P
from opentn_b200 import StiefelFit
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = P.kitaev_fullsites_target()
rng = np.random.default_rng(0)
t0=time.time()
# cheap synthetic inputs: random isometries via batched QR on GPU
g = torch.randn(B, 3, 256, 16, dtype=torch.float64, device='cuda')
q, r = torch.linalg.qr(g)
x = (q * torch.sign(torch.diagonal(r, dim1=-2, dim2=-1)).unsqueeze(-2)).contiguous()
z = torch.randn_like(x)
eta = (z - x @ (0.5 * (x.transpose(-1, -2) @ z + z.transpose(-1, -2) @ x))).contiguous()
print('inputs', time.time()-t0)
fit = StiefelFit(T, model='full', N=4, d=2, metric='canonical', max_batch=B)
f = torch.empty(B, dtype=torch.float64, device='cuda'); gr = torch.empty_like(x); hv = torch.empty_like(x)
for _ in range(3): fit.triple(x, eta, out=(f, gr, hv))
torch.cuda.synchronize()
for name, fn in [('triple', lambda: fit.triple(x, eta, out=(f, gr, hv))), ('cost', lambda: fit.cost(x)), ('cost_grad', lambda: fit.cost_grad(x))]:
    fn(); torch.cuda.synchronize()
    ts=[]
    for _ in range(5):
        t0=time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter()-t0)
    ms = min(ts)*1e3
    print(f'{name}: {ms:.2f} ms  -> {B/ms*1e3:.0f} evals/s')
fl = 0.642e9*B
print('triple TFLOP/s (algorithmic):', fl/ (min(ts) if False else 1) )
