"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.
Tolerances: cost and gradient 1e-10 relative (north star); in practice they agree to ~1e-13."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N, d = 4, 2


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def P():
    from oracle import port
    return port


@pytest.fixture(scope="module")
def targets(P):
    Lp = P.create_2local_liouvillians(P.pspl_lindbladians(1.0), N, d, pbc=True)[0]
    Lk = P.create_2local_liouvillians(P.get_kitaev_nn_linbladian(1.0), N, d, pbc=True)[0]
    return dict(pspl_tau1=P.exp_operator_dt(Lp, 1.0), pspl_tau05=P.exp_operator_dt(Lp, 0.5),
                kitaev_pbc_tau05=P.exp_operator_dt(Lk, 0.5), kitaev_obc_tau4=P.kitaev_fullsites_target())


def _random_batch(P, B, m, n, p, seed):
    rng = np.random.default_rng(seed)
    xs = np.array([[P.random_stiefel(n, p, rng) for _ in range(m)] for _ in range(B)])
    et = np.array([[P.project(x, rng.standard_normal((n, p))) for x in xb] for xb in xs])
    return xs, et


@pytest.mark.parametrize("dense", [False, True])
@pytest.mark.parametrize("metric", ["canonical", "euclidean"])
@pytest.mark.parametrize("m,K", [(3, 16), (3, 4), (1, 2), (2, 5)])
def test_full_model_triple(P, targets, metric, m, K, dense):
    from opentn_b200 import StiefelFit
    T = targets["kitaev_obc_tau4"]
    B = 3
    xs, et = _random_batch(P, B, m, 16 * K, 16, 7 + K)
    fit = StiefelFit(T, model="full", N=N, d=d, metric=metric, max_batch=B, dense=dense)
    f, g, h = fit.triple(xs, et)
    spec = P.CostSpec(T, "full")
    for b in range(B):
        fo, go, ho = P.triple(spec, list(xs[b]), list(et[b]), metric)
        assert abs(f[b] - fo) / fo < 1e-12
        assert rel(g[b], np.array(go)) < 1e-10
        assert rel(h[b], np.array(ho)) < 1e-10
    # cost-only and cost+grad entry points agree with the triple
    np.testing.assert_allclose(fit.cost(xs), f, rtol=1e-14)
    f2, g2, e2 = fit.cost_grad(xs, want_egrad=True)
    np.testing.assert_allclose(f2, f, rtol=1e-14)
    np.testing.assert_allclose(g2, g, rtol=0, atol=1e-14)
    for b in range(B):
        assert rel(e2[b], np.array(P.egrad(spec, list(xs[b])))) < 1e-10
    # cached HVP (what tCG uses) == triple's HVP
    h2 = fit.hvp_cached(et)
    np.testing.assert_allclose(h2, h, rtol=0, atol=1e-13)
    fit.close()


@pytest.mark.parametrize("dense", [False, True])
@pytest.mark.parametrize("metric", ["canonical", "euclidean"])
@pytest.mark.parametrize("m,K,tname", [(3, 10, "pspl_tau1"), (9, 16, "kitaev_pbc_tau05"), (5, 2, "pspl_tau05")])
def test_local_model_triple(P, targets, metric, m, K, tname, dense):
    from opentn_b200 import StiefelFit
    T = targets[tname]
    B = 2
    xs, et = _random_batch(P, B, m, 4 * K, 4, 11 + m)
    fit = StiefelFit(T, model="local", N=N, d=d, metric=metric, max_batch=B, dense=dense)
    f, g, h = fit.triple(xs, et)
    spec = P.CostSpec(T, "local", N, d)
    for b in range(B):
        fo, go, ho = P.triple(spec, list(xs[b]), list(et[b]), metric)
        assert abs(f[b] - fo) / fo < 1e-12
        assert rel(g[b], np.array(go)) < 1e-10
        assert rel(h[b], np.array(ho)) < 1e-10
    M = fit.model_matrix(xs)
    for b in range(B):
        np.testing.assert_allclose(M[b], P.model_stiefel_local(list(xs[b]), N, d), rtol=0, atol=1e-13)
    fit.close()


def test_golden_costs_on_gpu(P, targets, goldens):
    """Every cost anchor of the reference (notebook prints / .npy artefacts) through the CUDA path."""
    from opentn_b200 import StiefelFit
    fit = StiefelFit(targets["pspl_tau1"], model="local", N=N, d=d)
    assert abs(fit(list(goldens["ref/xs0_pspl_n1_K10"])) - goldens["known/pspl_n1_tau1_K10_cost0"]) < 1e-13
    assert abs(fit(goldens["art/xs_timestep_1_opt_4"]) - goldens["known/pspl_opt_cost_n1"]) < 1e-14
    assert abs(fit(goldens["art/xs_timestep_4_opt_9"]) - goldens["ref/cost_xs_timestep_4_opt_9"]) < 1e-14
    hist = [fit(x) for x in goldens["art/xs_timestep_1_opt_first8"]]
    np.testing.assert_allclose(hist, goldens["art/f_pauli_n1_tau1_rank10"][:8], rtol=1e-11)
    fit05 = StiefelFit(targets["pspl_tau05"], model="local", N=N, d=d)
    for n_ in (2, 3, 4):
        f = fit05(goldens[f"art/x_pauli_n{n_}_rank_10_tau05"])
        assert abs(f - goldens[f"ref/cost_x_pauli_n{n_}_rank_10_tau05"]) / f < 1e-11
    fk = StiefelFit(targets["kitaev_pbc_tau05"], model="local", N=N, d=d)
    f0 = fk(goldens["ref/xs0_kitaev_tau05_n4_K16"])
    assert abs(f0 - goldens["known/kitaev_n4_tau05_K16_cost0"]) / f0 < 1e-10   # cancellation-limited anchor (f ~ 5.7e-5)
    f1 = fk(goldens["art/x_kitaev_n4_rank_16"])
    assert abs(f1 - goldens["ref/cost_x_kitaev_n4_rank_16"]) / f1 < 1e-10
    ff = StiefelFit(targets["kitaev_obc_tau4"], model="full", N=N, d=d)
    xr4 = goldens["art/data_trust_region_stiefel_r4_xs"].real
    assert abs(ff(xr4) - goldens["ref/cost_fullsites_r4_xs"]) < 1e-13
    for x in (fit, fit05, fk, ff):
        x.close()


@pytest.mark.parametrize("model,m,K", [("full", 3, 8), ("local", 4, 3)])
def test_generic_complex_target(P, model, m, K):
    """A target that is NOT a physical superoperator (random, complex, no swap symmetry): the device path runs its reverse
    chain on the swap-symmetrised residual and folds the rest into a constant -- results must still equal the oracle's."""
    from opentn_b200 import StiefelFit
    rng = np.random.default_rng(99)
    T = 0.1 * (rng.standard_normal((256, 256)) + 1j * rng.standard_normal((256, 256)))
    p = 16 if model == "full" else 4
    xs, et = _random_batch(P, 2, m, p * K, p, 3)
    fit = StiefelFit(T, model=model, N=N, d=d, metric="canonical", max_batch=2, dense=False)
    f, g, h = fit.triple(xs, et)
    spec = P.CostSpec(T, model, N, d)
    for b in range(2):
        fo, go, ho = P.triple(spec, list(xs[b]), list(et[b]), "canonical")
        assert abs(f[b] - fo) / fo < 1e-12
        assert rel(g[b], np.array(go)) < 1e-10
        assert rel(h[b], np.array(ho)) < 1e-10
    fit.close()


def test_multiple_targets(P, targets):
    """Per-instance target selection (tau sweep): instance b fits target index[b]."""
    from opentn_b200 import StiefelFit
    Ts = np.stack([targets["pspl_tau1"], targets["pspl_tau05"], targets["kitaev_pbc_tau05"]])
    xs, et = _random_batch(P, 4, 3, 8, 4, 21)
    fit = StiefelFit(Ts, model="local", N=N, d=d, metric="canonical", max_batch=4, dense=False)
    index = [2, 0, 1, 0]
    fit.set_target_index(index)
    f, g, h = fit.triple(xs, et)
    for b in range(4):
        fo, go, ho = P.triple(P.CostSpec(Ts[index[b]], "local", N, d), list(xs[b]), list(et[b]), "canonical")
        assert abs(f[b] - fo) / fo < 1e-12
        assert rel(g[b], np.array(go)) < 1e-10
        assert rel(h[b], np.array(ho)) < 1e-10
    fit.close()


def test_pipelined_host_path_equals_device_path(P, targets):
    """otn_triple with host arrays and batch >= 512 takes the chunked H2D/compute/D2H pipeline; results must be bitwise
    those of the device-resident path (same kernels, same per-instance arithmetic)."""
    import torch
    from opentn_b200 import StiefelFit
    B, K = 600, 2
    rng = np.random.default_rng(1)
    g = rng.standard_normal((B, 3, 16 * K, 16))
    xs = np.linalg.qr(g)[0]
    z = rng.standard_normal(xs.shape)
    xtz = np.swapaxes(xs, -1, -2) @ z
    et = z - xs @ (0.5 * (xtz + np.swapaxes(xtz, -1, -2)))
    fit = StiefelFit(targets["kitaev_obc_tau4"], model="full", N=N, d=d, metric="canonical", max_batch=B)
    f_h, g_h, h_h = fit.triple(xs, et)
    xd, ed = torch.from_numpy(xs).cuda(), torch.from_numpy(et).cuda()
    f_d, g_d, h_d = fit.triple(xd, ed)
    np.testing.assert_array_equal(f_h, f_d.cpu().numpy())
    np.testing.assert_array_equal(g_h, g_d.cpu().numpy())
    np.testing.assert_array_equal(h_h, h_d.cpu().numpy())
    spec = P.CostSpec(targets["kitaev_obc_tau4"], "full")
    for b in (0, 255, 256, 599):
        fo, go, ho = P.triple(spec, list(xs[b]), list(et[b]), "canonical")
        assert abs(f_h[b] - fo) / fo < 1e-12 and rel(h_h[b], np.array(ho)) < 1e-10
    fit.close()


@pytest.mark.parametrize("model", ["local", "full"])
def test_pipelined_host_path_with_per_instance_targets(P, targets, model):
    """The chunked host pipeline must use each chunk's own slice of the target indices and of the workspaces (chunks overlap on
    two compute streams): per-instance targets, batch >= 512, both models (the 2-site model takes the factored evaluation)."""
    import torch
    from opentn_b200 import StiefelFit
    B, K = 640, 2
    p = 4 if model == "local" else 16
    rng = np.random.default_rng(2)
    xs = np.linalg.qr(rng.standard_normal((B, 3, p * K, p)))[0]
    z = rng.standard_normal(xs.shape)
    xtz = np.swapaxes(xs, -1, -2) @ z
    et = z - xs @ (0.5 * (xtz + np.swapaxes(xtz, -1, -2)))
    both = np.stack([targets["pspl_tau1"], targets["kitaev_pbc_tau05"]])
    idx = rng.integers(0, 2, B).astype(np.int32)
    fit = StiefelFit(both, model=model, N=N, d=d, metric="canonical", max_batch=B)
    fit.set_target_index(idx)
    f_h, g_h, h_h = fit.triple(xs, et)
    f_d, g_d, h_d = fit.triple(torch.from_numpy(xs).cuda(), torch.from_numpy(et).cuda())
    np.testing.assert_array_equal(f_h, f_d.cpu().numpy())
    np.testing.assert_array_equal(g_h, g_d.cpu().numpy())
    np.testing.assert_array_equal(h_h, h_d.cpu().numpy())
    for b in (0, 127, 128, 511, 512, 639):
        spec = P.CostSpec(both[idx[b]], model, N, d)
        fo, go, ho = P.triple(spec, list(xs[b]), list(et[b]), "canonical")
        assert abs(f_h[b] - fo) / fo < 1e-12 and rel(g_h[b], np.array(go)) < 1e-10 and rel(h_h[b], np.array(ho)) < 1e-10
    fit.close()


@pytest.mark.parametrize("metric", ["canonical", "euclidean"])
@pytest.mark.parametrize("m,K", [(3, 16), (2, 3), (1, 5)])
def test_dense_formulation_full_model(P, targets, metric, m, K):
    """OTN_FLAG_DENSE: the dense D x D chain (what the reference executes) against the oracle and against the default
    block-diagonal evaluation."""
    from opentn_b200 import StiefelFit
    T = targets["kitaev_obc_tau4"]
    B = 2
    xs, et = _random_batch(P, B, m, 16 * K, 16, 40 + K)
    dense = StiefelFit(T, model="full", N=N, d=d, metric=metric, max_batch=B, dense=True)
    blocks = StiefelFit(T, model="full", N=N, d=d, metric=metric, max_batch=B, dense=False)
    fd, gd, hd = dense.triple(xs, et)
    fb, gb, hb = blocks.triple(xs, et)
    np.testing.assert_allclose(fd, fb, rtol=1e-13)
    assert rel(gd, gb) < 1e-11 and rel(hd, hb) < 1e-11
    spec = P.CostSpec(T, "full")
    for b in range(B):
        fo, go, ho = P.triple(spec, list(xs[b]), list(et[b]), metric)
        assert abs(fd[b] - fo) / fo < 1e-12
        assert rel(gd[b], np.array(go)) < 1e-10 and rel(hd[b], np.array(ho)) < 1e-10
    Md, Mb = dense.model_matrix(xs), blocks.model_matrix(xs)
    for b in range(B):
        Mo = P.model_Ys_stiefel(list(xs[b]))
        np.testing.assert_allclose(Md[b], Mo, atol=1e-13)
        np.testing.assert_allclose(Mb[b], Mo, atol=1e-13)
    dense.close(); blocks.close()


def test_full_model_n3_dim8(P):
    """dim = 8 (N = 3 qubits, D = 64): exercises the DIM = 8 instantiations (blocks 36 -> 48 and 28 -> 64)."""
    from opentn_b200 import StiefelFit
    rng = np.random.default_rng(8)
    T = rng.standard_normal((64, 64))
    xs, et = _random_batch(P, 2, 3, 8 * 5, 8, 77)
    spec = P.CostSpec(T, "full")
    for dense in (False, True):
        fit = StiefelFit(T, model="full", N=3, d=2, metric="canonical", max_batch=2, dense=dense)
        f, g, h = fit.triple(xs, et)
        for b in range(2):
            fo, go, ho = P.triple(spec, list(xs[b]), list(et[b]), "canonical")
            assert abs(f[b] - fo) / fo < 1e-12
            assert rel(g[b], np.array(go)) < 1e-10 and rel(h[b], np.array(ho)) < 1e-10
        fit.close()
