"""Edge cases through the C ABI: large / ragged Kraus ranks, argument errors, degenerate trust-region inputs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
N, d = 4, 2


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


@pytest.mark.parametrize("K", [1, 7, 20, 24])
def test_ragged_and_large_kraus_rank(oracle, K):
    """K not a multiple of 4/8 (zero-padded Kraus rows in shared memory), the largest rank of the DMMA epilogue (20) and of the
    CUDA-core epilogue (24)."""
    from opentn_b200 import StiefelFit
    P = oracle
    T = P.kitaev_fullsites_target()
    rng = np.random.default_rng(K)
    xs = np.array([[P.random_stiefel(16 * K, 16, rng) for _ in range(2)]])
    et = np.array([[P.project(x, rng.standard_normal(x.shape)) for x in xs[0]]])
    spec = P.CostSpec(T, "full")
    fo, go, ho = P.triple(spec, list(xs[0]), list(et[0]), "canonical")
    for dense in (False, True):
        fit = StiefelFit(T, model="full", N=N, d=d, dense=dense)
        f, g, h = fit.triple(xs, et)
        assert abs(f[0] - fo) / fo < 1e-12 and rel(g[0], np.array(go)) < 1e-10 and rel(h[0], np.array(ho)) < 1e-10
        fit.close()


def test_kraus_rank_32(oracle, monkeypatch):
    """K = 32 (the assembly limit): cost, gradient and -- with the streaming Riemannian epilogue, which stages nothing n x p in
    shared memory -- the HVP work.  The staged epilogue (OTN_RIEM_STAGED=1: four n x p arrays resident in shared memory) exceeds
    the 227 KB budget there and must say so (no silent fallback)."""
    from opentn_b200 import OtnError, StiefelFit
    P = oracle
    T = P.kitaev_fullsites_target()
    rng = np.random.default_rng(32)
    xs = np.array([[P.random_stiefel(16 * 32, 16, rng) for _ in range(2)]])
    spec = P.CostSpec(T, "full")
    fit = StiefelFit(T, model="full", N=N, d=d)
    f = fit.cost(xs)
    assert abs(f[0] - spec(list(xs[0]))) / f[0] < 1e-12
    f2, g = fit.cost_grad(xs)
    go = P.rgrad(spec, list(xs[0]), "canonical")
    assert rel(g[0], np.array(go)) < 1e-10
    f3, g3, h3 = fit.triple(xs, g)
    fo, go, ho = P.triple(spec, list(xs[0]), list(g[0]), "canonical")
    assert abs(f3[0] - fo) / fo < 1e-12 and rel(g3[0], np.array(go)) < 1e-10 and rel(h3[0], np.array(ho)) < 1e-10
    fit.close()
    monkeypatch.setenv("OTN_RIEM_STAGED", "1")
    fit = StiefelFit(T, model="full", N=N, d=d)
    with pytest.raises(OtnError, match="Kraus rank <= 24"):
        fit.triple(xs, g)
    fit.close()


def test_argument_errors():
    from opentn_b200 import OtnError, StiefelFit
    fit = StiefelFit(np.eye(256), model="full", N=N, d=d, max_batch=2)
    x = np.linalg.qr(np.random.default_rng(0).standard_normal((2, 3, 32, 16)))[0]
    with pytest.raises(OtnError, match="no primal cache"):
        fit._cached = (fit._handle(3, 2, 2), 2, None, False)
        fit.hvp_cached(x)
    with pytest.raises(ValueError, match="does not fit model"):
        fit.cost(np.zeros((1, 3, 32, 8)))
    with pytest.raises(OtnError, match="Kraus rank"):
        fit.cost(np.zeros((1, 1, 16 * 33, 16)))
    fit.close()
    with pytest.raises(OtnError, match="Only even number of sites"):          # transformations.py:414
        StiefelFit(np.eye(64), model="local", N=3, d=2).cost(np.zeros((1, 1, 8, 4)))


def test_tr_flags_broken_instance_without_aborting_the_batch(oracle):
    """One instance of the batch is numerically broken (a NaN in its isometry): the batched solver flags it in `status`
    (OTN_STATUS_NAN), never moves it, and the other instance is fitted as if it were alone ('never abort the batch')."""
    from opentn_b200 import StiefelFit
    P = oracle
    T = P.kitaev_fullsites_target()
    rng = np.random.default_rng(2)
    xs = np.array([[P.random_stiefel(32, 16, rng) for _ in range(2)] for _ in range(2)])
    bad = xs.copy()
    bad[0, 1, 5, 3] = np.nan
    fit = StiefelFit(T, model="full", N=N, d=d, max_batch=2)
    x, rep = fit.tr_run(bad, niter=2)
    assert rep["status"][0] & 1                       # OTN_STATUS_NAN
    assert rep["status"][1] == 0
    assert (rep["accepted"][:, 0] == 0).all()
    good, rep1 = fit.tr_run(xs[1:2], niter=2)
    np.testing.assert_array_equal(rep["f_iter"][:, 1], rep1["f_iter"][:, 0])
    np.testing.assert_array_equal(x[1], good[0])
    # f = 0 up to rounding (target == model(x)): finite garbage direction, but no NaN and no crash
    fit0 = StiefelFit(P.model_Ys_stiefel(list(xs[0])), model="full", N=N, d=d)
    f0 = fit0.cost(xs[0:1])
    assert 0 <= f0[0] < 1e-13
    x0, rep0 = fit0.tr_run(xs[0:1], niter=1)
    assert np.isfinite(x0).all()
    fit.close(); fit0.close()


def test_tr_zero_iterations_and_radius_restart(oracle):
    from opentn_b200 import StiefelFit
    P = oracle
    T = P.kitaev_fullsites_target()
    rng = np.random.default_rng(3)
    xs = np.array([[P.random_stiefel(32, 16, rng) for _ in range(3)]])
    fit = StiefelFit(T, model="full", N=N, d=d)
    x0, rep0 = fit.tr_run(xs, niter=0)
    np.testing.assert_array_equal(x0, xs)
    # warm restart (trust_region_rcopt.py:37): 2 + 2 iterations with the radius carried over == 4 iterations
    xa, ra = fit.tr_run(xs, niter=2)
    xb, rb = fit.tr_run(xa, niter=2, radius=ra["radius"])
    xc, rc = fit.tr_run(xs, niter=4)
    np.testing.assert_allclose(np.concatenate([ra["f_iter"], rb["f_iter"]])[:, 0], rc["f_iter"][:, 0], rtol=1e-12)
    np.testing.assert_allclose(xb, xc, atol=1e-12)
    fit.close()
