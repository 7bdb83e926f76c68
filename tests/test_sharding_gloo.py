"""World-size-2 gloo test (CPU) of the N>1 path: round-robin instance sharding + the single final all-gather."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from opentn_b200 import sharding
    from oracle import port as P
    # tiny full-sites problem (N=2: D=16, isometries (8,4)); the "fit" of the CPU test is the oracle cost + one retraction
    T = P.exp_operator_dt(P.lindbladian2super([P.get_kitaev_nn_linbladian(1.0)]), 1.0)
    spec = P.CostSpec(T, "full")
    rng = np.random.default_rng(0)                      # same inputs on every rank
    xs_all = torch.from_numpy(np.array([[P.random_stiefel(8, 4, rng) for _ in range(2)] for _ in range(total)]))

    def run_local(x):
        xn = x.numpy()
        f = torch.tensor([spec(list(xb)) for xb in xn])
        return x * 1.0, f, torch.zeros(len(xn), dtype=torch.int32)

    x_all, f_all, s_all = sharding.sharded_fit(run_local, xs_all)
    idx = sharding.my_instances(total, rank, world)
    assert list(idx) == list(range(rank, total, world))
    want = torch.tensor([spec(list(xb)) for xb in xs_all.numpy()])
    assert torch.equal(f_all, want), (f_all, want)
    assert torch.equal(x_all, xs_all)
    assert s_all.shape == (total,)
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")


@pytest.mark.parametrize("total", [7, 8])
def test_sharded_gather_world2(tmp_path, total):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def test_shard_sizes():
    from opentn_b200 import sharding
    assert sharding.shard_sizes(4096, 8) == [512] * 8
    assert sharding.shard_sizes(7, 2) == [4, 3]
    assert sum(sharding.shard_sizes(1000, 8)) == 1000
    x = torch.arange(5.0)
    assert torch.equal(sharding.gather_instances(x, 5), x)       # no process group: identity


# ---- config-4 sweep: cell partition + group-wise final gather (opentn_b200.sweep) ---------------------------------------------
def _cells_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from opentn_b200 import sweep
    ranks, n_taus, n_seeds, m = [2, 4, 8], 2, 3, 2                     # 6 cells: 3 per rank, every K is split between the ranks
    cells = sweep.plan_cells(ranks, n_taus, world, rank)
    res = {}
    for K in sorted({k for k, _ in cells}):
        tis = [ti for k, ti in cells if k == K]
        B = len(tis) * n_seeds
        tindex = torch.tensor(np.repeat(tis, n_seeds), dtype=torch.int32)
        seed = torch.tensor(np.tile(np.arange(n_seeds), len(tis)))
        f0 = (1000.0 * K + 10.0 * tindex + seed).to(torch.float64)       # a value that identifies the instance
        res[K] = dict(tau_index=tindex, seed=seed, f0=f0, x0=f0[:, None, None, None] * torch.ones(B, m, 4 * K, 4, dtype=torch.float64))
    full = sweep.gather_groups(res, ranks, n_taus, n_seeds, world)
    for K in ranks:
        r = full[K]
        want = torch.tensor([1000.0 * K + 10.0 * ti + s for ti in range(n_taus) for s in range(n_seeds)], dtype=torch.float64)
        assert torch.equal(r["f0"], want), (K, r["f0"], want)
        assert r["x0"].shape == (n_taus * n_seeds, m, 4 * K, 4) and torch.equal(r["x0"][:, 0, 0, 0], want)
        assert r["tau_index"].tolist() == [ti for ti in range(n_taus) for _ in range(n_seeds)]
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")


def test_sweep_cells_gather_world2(tmp_path):
    world = 2
    mp.spawn(_cells_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def test_plan_cells_partition():
    from opentn_b200 import sweep
    ranks, n_taus = [2, 4, 8, 16], 4
    for world in (1, 2, 4, 8):
        owned = [sweep.plan_cells(ranks, n_taus, world, r) for r in range(world)]
        assert sorted(sum(owned, [])) == sorted((K, ti) for K in ranks for ti in range(n_taus))       # a partition
        assert {len(o) for o in owned} == {16 // world}
        if world > 1:      # snake deal: every rank's cells pair small isometries with large ones
            assert all(min(K for K, _ in o) <= 4 and max(K for K, _ in o) >= 8 for o in owned)
    assert sweep.instance_seed(16, 3, 255) != sweep.instance_seed(16, 2, 255)


def _row_shard_worker(rank, world, port, out_dir):
    """Host logic of StiefelFit.row_shard under a world-size-2 gloo group: rank / world bookkeeping and the divisibility check
    (D = 256 over 2 ranks is fine; D = 64 is not: 64 % (64 * 2) != 0).  No handle is created, so no GPU is needed."""
    import os
    import numpy as np
    import pytest
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from opentn_b200 import StiefelFit
    fit = StiefelFit(np.eye(256), model="local", N=4, d=2, dense=True).row_shard()
    assert fit._row_shard[:2] == (rank, world) and fit._handles == {}
    with pytest.raises(ValueError, match="multiple of 64"):
        StiefelFit(np.eye(64), model="full", N=3, d=2, dense=True).row_shard()
    fit.close()                                   # nothing attached: no collective
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(out_dir, f"rs_ok{rank}"), "w").write("ok")


def test_row_shard_host_logic_two_ranks(tmp_path):
    import os
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_row_shard_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all(os.path.exists(tmp_path / f"rs_ok{r}") for r in range(2))
