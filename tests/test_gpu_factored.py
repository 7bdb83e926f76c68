"""Factored (never-materialise-Y_full) evaluation of the 2-site model, OTN_FLAG_FACTORED (SURVEY 8f "next" #3), through the
C ABI against the CPU oracle and against the default (block-diagonal) device path.  FP64: cost 1e-12, gradient / HVP 1e-10
relative, as for the other formulations."""
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
    return dict(pspl_tau1=P.exp_operator_dt(Lp, 1.0), kitaev_pbc_tau05=P.exp_operator_dt(Lk, 0.5))


def _random_batch(P, B, m, n, p, seed):
    rng = np.random.default_rng(seed)
    xs = np.array([[P.random_stiefel(n, p, rng) for _ in range(m)] for _ in range(B)])
    et = np.array([[P.project(x, rng.standard_normal((n, p))) for x in xb] for xb in xs])
    return xs, et


@pytest.mark.parametrize("max_batch", [3, 64])      # <= 4 instances: 2-column tiles; more: 8-column tiles (two kernel variants)
@pytest.mark.parametrize("metric", ["canonical", "euclidean"])
@pytest.mark.parametrize("m,K,tname", [(3, 10, "pspl_tau1"), (9, 16, "kitaev_pbc_tau05"), (1, 3, "pspl_tau1"), (4, 2, "pspl_tau1")])
def test_factored_triple_vs_oracle(P, targets, metric, m, K, tname, max_batch):
    from opentn_b200 import StiefelFit
    T = targets[tname]
    B = 3
    xs, et = _random_batch(P, B, m, 4 * K, 4, 21 + m)
    fit = StiefelFit(T, model="local", N=N, d=d, metric=metric, max_batch=max_batch, factored=True)
    M = fit.model_matrix(xs)
    for b in range(B):
        np.testing.assert_allclose(M[b], P.model_stiefel_local(list(xs[b]), N, d), rtol=0, atol=1e-13)
    f, g, h = fit.triple(xs, et)
    spec = P.CostSpec(T, "local", N, d)
    for b in range(B):
        fo, go, ho = P.triple(spec, list(xs[b]), list(et[b]), metric)
        assert abs(f[b] - fo) / fo < 1e-12
        assert rel(g[b], np.array(go)) < 1e-10
        assert rel(h[b], np.array(ho)) < 1e-10
    np.testing.assert_allclose(fit.cost(xs), f, rtol=1e-14)
    f2, g2 = fit.cost_grad(xs)
    np.testing.assert_allclose(f2, f, rtol=1e-14)
    np.testing.assert_allclose(g2, g, rtol=0, atol=1e-13)
    np.testing.assert_allclose(fit.hvp_cached(et), h, rtol=0, atol=1e-12)
    fit.close()


def test_factored_generic_complex_target(P):
    """Non-physical complex target: only its swap-symmetric real part carries gradient; the rest is a constant."""
    from opentn_b200 import StiefelFit
    rng = np.random.default_rng(99)
    T = 0.1 * (rng.standard_normal((256, 256)) + 1j * rng.standard_normal((256, 256)))
    xs, et = _random_batch(P, 2, 4, 12, 4, 3)
    fit = StiefelFit(T, model="local", N=N, d=d, max_batch=2, factored=True)
    f, g, h = fit.triple(xs, et)
    spec = P.CostSpec(T, "local", N, d)
    for b in range(2):
        fo, go, ho = P.triple(spec, list(xs[b]), list(et[b]), "canonical")
        assert abs(f[b] - fo) / fo < 1e-12
        assert rel(g[b], np.array(go)) < 1e-10
        assert rel(h[b], np.array(ho)) < 1e-10
    fit.close()


def test_factored_per_instance_targets_and_golden_costs(P, targets, goldens):
    from opentn_b200 import StiefelFit
    fit = StiefelFit(targets["pspl_tau1"], model="local", N=N, d=d, factored=True)
    assert abs(fit(list(goldens["ref/xs0_pspl_n1_K10"])) - goldens["known/pspl_n1_tau1_K10_cost0"]) < 1e-13
    assert abs(fit(goldens["art/xs_timestep_1_opt_4"]) - goldens["known/pspl_opt_cost_n1"]) < 1e-14
    fit.close()
    both = np.stack([targets["pspl_tau1"], targets["kitaev_pbc_tau05"]])
    xs, _ = _random_batch(P, 4, 3, 8, 4, 5)
    fit = StiefelFit(both, model="local", N=N, d=d, max_batch=4, factored=True)
    idx = [1, 0, 0, 1]
    fit.set_target_index(idx)
    f = fit.cost(xs)
    for b in range(4):
        fo = P.cost_and_egrad(P.CostSpec(both[idx[b]], "local", N, d), list(xs[b]))[0]
        assert abs(f[b] - fo) / fo < 1e-12
    fit.close()


def test_factored_trust_region_matches_default_path(P, targets, goldens):
    """The batched device solver on the factored evaluation follows the same trajectory as on the block-diagonal one
    (same algorithm, results equal to rounding for the first iterations) and reproduces the reference's stored history."""
    from opentn_b200 import StiefelFit
    xs0 = np.array([goldens["ref/xs0_pspl_n1_K10"]])
    out = {}
    for name, kw in (("factored", dict(factored=True)), ("blocks", dict(dense=False))):
        fit = StiefelFit(targets["pspl_tau1"], model="local", N=N, d=d, **kw)
        x, rep = fit.tr_run(xs0, niter=5)
        out[name] = (x, rep)
        fit.close()
    np.testing.assert_allclose(out["factored"][1]["f_iter"], out["blocks"][1]["f_iter"], rtol=1e-9)
    np.testing.assert_allclose(out["factored"][1]["f_iter"][:5, 0], goldens["art/f_pauli_n1_tau1_rank10"][:5], rtol=1e-8)


@pytest.mark.slow
def test_factored_n6(oracle, goldens):
    """N = 6 (D = 4096, three factors per layer, one column per CTA): the reference's N=6 anchors and the block-diagonal path."""
    from opentn_b200 import StiefelFit
    from opentn_b200 import stiefel as S
    from opentn_b200 import transformations as T
    P = oracle
    Lnn = P.get_kitaev_nn_linbladian(1.0)
    L6 = T.create_2local_liouvillians(Lnn, 6, 2, pbc=True, backend="cuda")[0]
    T6 = T.exp_operator_dt(L6, 4.0, backend="cuda")
    xs = [P.super2ortho(x.real, rank=2) for x in P.get_general_trotter_local_ansatz([Lnn], 4, n=1)]
    fit = StiefelFit(T6, model="local", N=6, d=2, metric="canonical", factored=True)
    f, g = fit.cost_grad(xs)
    assert abs(f - 0.5394470932876375) < 1e-11
    gn = np.sqrt(sum(S.canonical_metric(gl, gl, x) for gl, x in zip(g, xs)))
    assert abs(gn - goldens["known/n6_gradnorm"]) / goldens["known/n6_gradnorm"] < 1e-10
    rng = np.random.default_rng(0)
    v = np.array([S.project(x, rng.standard_normal(x.shape)) for x in xs])
    ref = StiefelFit(T6, model="local", N=6, d=2, metric="canonical", dense=False)
    f1, g1, h1 = ref.triple(xs, v)
    f2, g2, h2 = fit.triple(xs, v)
    assert abs(f1 - f2) < 1e-12
    assert rel(g2, g1) < 1e-10
    assert rel(h2, h1) < 1e-10
    ref.close()
    fit.close()


@pytest.mark.parametrize("m,K", [(2, 1), (17, 2), (6, 7)])
def test_factored_edge_shapes(P, targets, m, K):
    """Square isometries (K = 1), an even / long layer stack, an odd Kraus rank: factored == oracle and == block-diagonal."""
    from opentn_b200 import StiefelFit
    T = targets["pspl_tau1"]
    B = 2
    xs, et = _random_batch(P, B, m, 4 * K, 4, 100 + m)
    fit = StiefelFit(T, model="local", N=N, d=d, max_batch=16, factored=True)
    ref = StiefelFit(T, model="local", N=N, d=d, max_batch=B, dense=False)
    f, g, h = fit.triple(xs, et)
    f2, g2, h2 = ref.triple(xs, et)
    np.testing.assert_allclose(f, f2, rtol=1e-13)
    assert rel(g, g2) < 1e-11 and rel(h, h2) < 1e-11
    fo, go, ho = P.triple(P.CostSpec(T, "local", N, d), list(xs[0]), list(et[0]), "canonical")
    assert abs(f[0] - fo) / fo < 1e-12 and rel(g[0], np.array(go)) < 1e-10 and rel(h[0], np.array(ho)) < 1e-10
    fit.close(); ref.close()


def test_factored_flag_validation(targets):
    """OTN_FLAG_FACTORED is refused where it does not apply (full-sites model, d = 3, N = 2) and cannot be combined with the
    other formulations; the default picks it only for the 2-site model with d = 2 and N in {4, 6}."""
    import ctypes as C
    from opentn_b200 import _lib
    lib = _lib.load()
    T = np.ascontiguousarray(targets["pspl_tau1"].real)
    h = C.c_void_p()
    # (N, d, model, flags) -> expected failure
    for Nn, dd, model, flags in [(4, 2, _lib.MODEL_FULL, 4), (4, 2, _lib.MODEL_LOCAL, 4 | 1), (4, 2, _lib.MODEL_LOCAL, 4 | 2)]:
        rc = lib.otn_create_ex(C.byref(h), Nn, dd, 3, 2, model, 1, 1, 1, T.ctypes.data, None, 0, flags)
        assert rc != 0 and b"FACTORED" in lib.otn_last_error()


def test_factored_batched_tr_vs_oracle(P, targets):
    """Batched device TR on the factored evaluation against the oracle's TR loop, instance by instance (tCG lengths differ per
    instance, so the per-instance masks of the lock-step solver are exercised)."""
    from opentn_b200 import StiefelFit
    T = targets["kitaev_pbc_tau05"]
    B, m, K = 3, 3, 4
    xs, _ = _random_batch(P, B, m, 4 * K, 4, 77)
    fit = StiefelFit(T, model="local", N=N, d=d, max_batch=8, factored=True)
    x, rep = fit.tr_run(xs, niter=3)
    spec = P.CostSpec(T, "local", N, d)
    for b in range(B):
        _, f_it, _, _ = P.trust_region_fit(spec, list(xs[b]), "canonical", niter=3)
        np.testing.assert_allclose(rep["f_iter"][:, b], f_it, rtol=1e-9)
    fit.close()
