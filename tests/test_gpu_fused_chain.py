"""GPU parity of the fused per-instance chain kernel (csrc/chain_fused.cuh) against the CPU oracle and against the tiled
per-product kernels it replaces.  The fused kernel is what large batches of the block-diagonal evaluation run (headline
workload); `OTN_FUSED_CHAIN=1` forces it for the small batches used here, `=0` forces the tiled kernels."""
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


def _random_batch(P, B, m, n, p, seed):
    rng = np.random.default_rng(seed)
    xs = np.array([[P.random_stiefel(n, p, rng) for _ in range(m)] for _ in range(B)])
    et = np.array([[P.project(x, rng.standard_normal((n, p))) for x in xb] for xb in xs])
    return xs, et


def _fit(monkeypatch, fused, *a, **kw):
    from opentn_b200 import StiefelFit
    monkeypatch.setenv("OTN_FUSED_CHAIN", "1" if fused else "0")
    return StiefelFit(*a, **kw)


@pytest.mark.parametrize("metric", ["canonical", "euclidean"])
@pytest.mark.parametrize("m,K", [(3, 16), (2, 5), (4, 3), (5, 2), (9, 2)])
def test_fused_full_model_vs_oracle(P, monkeypatch, metric, m, K):
    T = P.kitaev_fullsites_target()
    B = 3
    xs, et = _random_batch(P, B, m, 16 * K, 16, 100 + 7 * m + K)
    fit = _fit(monkeypatch, True, T, model="full", N=N, d=d, metric=metric, max_batch=B, dense=False)
    f, g, h = fit.triple(xs, et)
    spec = P.CostSpec(T, "full")
    for b in range(B):
        fo, go, ho = P.triple(spec, list(xs[b]), list(et[b]), metric)
        assert abs(f[b] - fo) / fo < 1e-12
        assert rel(g[b], np.array(go)) < 1e-10
        assert rel(h[b], np.array(ho)) < 1e-10
    # every op program of the fused kernel: cost only (FWD), cost + gradient (FWD, REV), cached HVP (TANGENT), model (FWD_MODEL)
    np.testing.assert_allclose(fit.cost(xs), f, rtol=1e-14)
    f2, g2, e2 = fit.cost_grad(xs, want_egrad=True)
    np.testing.assert_allclose(f2, f, rtol=1e-14)
    np.testing.assert_allclose(g2, g, rtol=0, atol=1e-13)
    assert rel(e2[0], np.array(P.egrad(spec, list(xs[0])))) < 1e-10
    np.testing.assert_allclose(fit.hvp_cached(et), h, rtol=0, atol=1e-12)
    M = fit.model_matrix(xs[:1])
    assert rel(M[0], P.model_Ys_stiefel(list(xs[0]))) < 1e-12
    fit.close()


@pytest.mark.parametrize("m,K", [(3, 16), (4, 4)])
def test_fused_equals_tiled(P, monkeypatch, m, K):
    """Same contraction order per output element => the matrices agree to rounding of the cost partial sums only."""
    T = P.kitaev_fullsites_target()
    B = 5
    xs, et = _random_batch(P, B, m, 16 * K, 16, 55 + K)
    fa = _fit(monkeypatch, True, T, model="full", N=N, d=d, metric="canonical", max_batch=B, dense=False)
    ra = fa.triple(xs, et)
    fb = _fit(monkeypatch, False, T, model="full", N=N, d=d, metric="canonical", max_batch=B, dense=False)
    rb = fb.triple(xs, et)
    np.testing.assert_allclose(ra[0], rb[0], rtol=1e-14)
    assert rel(ra[1], rb[1]) < 1e-13 and rel(ra[2], rb[2]) < 1e-12
    fa.close(); fb.close()


def test_fused_local_blocks_vs_oracle(P, monkeypatch):
    """2-site model through the block-diagonal chain (dense=False) with the fused kernel."""
    Lk = P.create_2local_liouvillians(P.get_kitaev_nn_linbladian(1.0), N, d, pbc=True)[0]
    T = P.exp_operator_dt(Lk, 0.5)
    m, K, B = 5, 3, 2
    xs, et = _random_batch(P, B, m, 4 * K, 4, 77)
    fit = _fit(monkeypatch, True, T, model="local", N=N, d=d, metric="canonical", max_batch=B, dense=False)
    f, g, h = fit.triple(xs, et)
    spec = P.CostSpec(T, "local", N, d)
    for b in range(B):
        fo, go, ho = P.triple(spec, list(xs[b]), list(et[b]), "canonical")
        assert abs(f[b] - fo) / fo < 1e-12
        assert rel(g[b], np.array(go)) < 1e-10
        assert rel(h[b], np.array(ho)) < 1e-10
    fit.close()


def test_fused_trust_region_vs_oracle(P, monkeypatch):
    """Batched TR over the fused kernels (cached-forward path, masks for finished tCG instances) against the oracle loop."""
    T = P.kitaev_fullsites_target()
    m, K, B = 3, 4, 3
    xs, _ = _random_batch(P, B, m, 16 * K, 16, 31)
    fit = _fit(monkeypatch, True, T, model="full", N=N, d=d, metric="canonical", max_batch=B, dense=False)
    x, rep = fit.tr_run(xs, niter=3)
    spec = P.CostSpec(T, "full")
    for b in range(B):
        xs_o, f_o, rad, log_ = P.trust_region_fit(spec, list(xs[b]), "canonical", niter=3)
        np.testing.assert_allclose(rep["f_iter"][:, b], f_o, rtol=1e-9)
        assert rel(x[b], np.array(xs_o[-1])) < 1e-7
    fit.close()


def test_fused_masked_instances(P, monkeypatch):
    """Per-instance targets and a large batch (one CTA per instance and block): every instance equals its single evaluation."""
    T0 = P.kitaev_fullsites_target()
    T1 = P.kitaev_fullsites_target(tau=1.0)
    m, K, B = 3, 2, 70
    base_x, base_e = _random_batch(P, 4, m, 16 * K, 16, 9)
    idx = np.arange(B) % 4
    xs, et = base_x[idx].copy(), base_e[idx].copy()
    fit = _fit(monkeypatch, True, np.stack([T0, T1]), model="full", N=N, d=d, metric="canonical", max_batch=B, dense=False)
    tix = (np.arange(B) // 4) % 2
    fit.set_target_index(tix)
    f, g, h = fit.triple(xs, et)
    for b in (0, 5, 37, 69):
        spec = P.CostSpec([T0, T1][tix[b]], "full")
        fo, go, ho = P.triple(spec, list(xs[b]), list(et[b]), "canonical")
        assert abs(f[b] - fo) / fo < 1e-12 and rel(g[b], np.array(go)) < 1e-10 and rel(h[b], np.array(ho)) < 1e-10
    # identical inputs + identical target => bitwise identical outputs wherever they sit in the batch
    same = [b for b in range(B) if idx[b] == idx[0] and tix[b] == tix[0]]
    for b in same[1:]:
        np.testing.assert_array_equal(g[b], g[same[0]])
        np.testing.assert_array_equal(h[b], h[same[0]])
    fit.close()
