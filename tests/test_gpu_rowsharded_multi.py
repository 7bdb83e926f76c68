"""Two-rank GPU test of the row-sharded composition (BASELINE config 5): both exchange modes -- NCCL all-gather after each
product, and the all-gather fused into the GEMM epilogue (NVLink peer stores, device-side arrival flags) -- against the ORACLE's
`compose_superops_list` / cost / layer cotangents (not against torch).  Skipped on a box with fewer than 2 GPUs."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from opentn_b200.rowsharded import RowShardedChain
    from oracle import port as P
    rng = np.random.default_rng(3)                      # same problem on every rank
    m = 4
    xs = [P.random_stiefel(12, 4, rng) for _ in range(m)]
    T = P.exp_operator_dt(P.create_2local_liouvillians(P.get_kitaev_nn_linbladian(1.0), 4, 2, pbc=True)[0], 0.5)
    layers = P.local_layers_full(xs, 4, 2)
    Ys = [torch.from_numpy(np.ascontiguousarray(y)).cuda() for y in layers]
    Tre = torch.from_numpy(np.ascontiguousarray(T.real)).cuda()
    spec = P.CostSpec(T, "local", 4, 2)
    want_M = P.compose_superops_list(layers)
    want_f = spec(xs)
    Z = P.egrad(spec, xs)
    for p2p in (False, True):
        ch = RowShardedChain(256, p2p=p2p, nbuf=3 * (m - 1))
        assert ch.world == world and ch.rows == 256 // world
        for rep in range(3):                            # repeated: the flag epochs and the buffer ring must stay consistent
            M = ch.compose(Ys)
            np.testing.assert_allclose(M.cpu().numpy(), want_M, rtol=0, atol=1e-13)
        assert abs(ch.cost(Ys, Tre) - want_f) < 1e-12
        f, W = ch.cost_and_layer_cotangents(Ys, Tre)
        assert abs(f - want_f) < 1e-12
        for l in range(m):
            z = P._layer_vjp(spec, l, xs[l], W[l].cpu().numpy())
            assert np.abs(z - Z[l]).max() / np.abs(Z[l]).max() < 1e-10
        # complex operands (targets / layers stored complex128 in the reference): (re, im) pairs
        A = rng.standard_normal((256, 256)) + 1j * rng.standard_normal((256, 256))
        B = rng.standard_normal((256, 256)) + 1j * rng.standard_normal((256, 256))
        tc = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()   # noqa: E731
        Cr, Ci = ch.matmul_complex((tc(A.real), tc(A.imag)), (tc(B.real), tc(B.imag)))
        np.testing.assert_allclose(Cr.cpu().numpy() + 1j * Ci.cpu().numpy(), P.compose_superops_list([B, A]), rtol=0, atol=1e-11)
        if p2p:
            with pytest.raises(ValueError, match="nbuf"):
                RowShardedChain.cost_and_layer_cotangents(ch, Ys + Ys, Tre)     # 8 layers need 21 ring buffers: refused, not corrupted
        ch.close()
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")


def test_rowsharded_two_ranks_vs_oracle(tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))
