"""Config 5: single N=6 instance (D=4096), the chain products row-sharded over the ranks (torchrun).  Prints per-rank-0 timings.
    python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/n6_rowsharded.py
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
from opentn_b200.rowsharded import RowShardedChain
from opentn_b200 import transformations as T
T.set_device(lr)

D, m = 4096, 3
torch.manual_seed(0)
# synthetic CPTP-like layers: product structure does not matter for the composition kernel; use random local layers embedded
rng = np.random.default_rng(0)
Ys = []
for l in range(m):
    q, _ = np.linalg.qr(rng.standard_normal((64, 4)))
    yl = T.ortho2super(q)
    Ys.append(torch.from_numpy(T.local2full(yl, 6, 2, shift_pbc=bool(l % 2))).cuda())
Tre = torch.randn(D, D, dtype=torch.float64, device="cuda") * 0.01
P2P = os.environ.get('OTN_P2P', '0') == '1'
ch = RowShardedChain(D, p2p=P2P, nbuf=8)
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): r = fn()
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / n], device="cuda")
    if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms), r
ms_c, M = timeit(lambda: ch.compose(Ys))
ms_f, f = timeit(lambda: ch.cost(Ys, Tre))
ms_g, fw = timeit(lambda: ch.cost_and_layer_cotangents(Ys, Tre))
Ysc = [(y, torch.zeros_like(y)) for y in Ys]
ms_z, Mz = timeit(lambda: ch.compose_complex(Ysc), n=2)
if rank == 0:
    ref = Ys[2] @ (Ys[1] @ Ys[0])
    err = float((M - ref).abs().max())
    fl = 2 * D ** 3
    print(f"world={world} p2p={P2P}: compose (2 products) {ms_c:.2f} ms = {2*fl/ms_c/1e9:.1f} TFLOP/s aggregate, max err vs torch {err:.2e}")
    print(f"          cost (2 products, last sharded) {ms_f:.2f} ms ; cost + layer cotangents (6 products) {ms_g:.2f} ms = {6*fl/ms_g/1e9:.1f} TFLOP/s")
    print(f"          complex compose (2 complex = 8 real products) {ms_z:.2f} ms = {8*fl/ms_z/1e9:.1f} TFLOP/s ; f = {f:.6f}")
ch.close()
if world > 1:
    dist.destroy_process_group()
