"""Parity at BASELINE.json's FULL sizes, through size-independent properties (the oracle takes ~25 ms per instance, so only a
sample of the 1024 instances is compared value by value; everything else is checked for all instances at once):

* config 3 (headline): Kitaev N=4 full-sites ansatz, m=3, K=16, B=1024 random-init instances (bench.make_inputs);
* config 4 shape     : Kitaev pbc tau=0.5 2-site ansatz, m=9, K=16, B=1024.

Properties: tangency of rgrad / H eta, Kraus-mixing invariance f((1 (x) U) x) = f(x) with covariant derivatives (the
superoperator sum_k E_k (x) E_k* does not see an orthogonal mixing of the Kraus index, transformations.py:743-749), symmetry and
linearity of the Riemannian Hessian in the metric (stiefel.py:477-531 builds a symmetric matrix from it), first / second
directional derivatives of f along the polar retraction (a second-order retraction, stiefel.py:323-355), batch independence,
agreement of the three formulations (dense, block-diagonal, factored), retraction onto the manifold, TR monotonicity."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N, d = 4, 2
B_FULL = 1024


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def cdot(x, a, b):
    """canonical inner product per instance, sum over layers of tr(a^T (1 - x x^T / 2) b) (stiefel.py:76-83)"""
    xa = np.einsum("blnp,blnq->blpq", x, a)
    xb = np.einsum("blnp,blnq->blpq", x, b)
    return np.einsum("blnp,blnp->b", a, b) - 0.5 * np.einsum("blpq,blpq->b", xa, xb)


def sym_part(x, z):
    s = np.einsum("blnp,blnq->blpq", x, z)
    return s + s.transpose(0, 1, 3, 2)


def random_orthogonal(k, rng):
    q, r = np.linalg.qr(rng.standard_normal((k, k)))
    return q * np.sign(np.diag(r))


def mix_kraus(x, U):
    """rows are (a, k) = a*K + k (transformations.py:716-726): x'[(a,k),c] = sum_k' U[l][k,k'] x[(a,k'),c], one U per layer"""
    Bq, m, n, p = x.shape
    K = U.shape[-1]
    return np.einsum("lkj,blajc->blakc", U, x.reshape(Bq, m, n // K, K, p)).reshape(Bq, m, n, p)


@pytest.fixture(scope="module")
def P():
    from oracle import port
    return port


@pytest.fixture(scope="module")
def c3(P):
    import bench
    xs, et = bench.make_inputs(0, B_FULL)
    return P.kitaev_fullsites_target(), xs, et


@pytest.fixture(scope="module")
def c3_triple(c3):
    from opentn_b200 import StiefelFit
    T, xs, et = c3
    fit = StiefelFit(T, model="full", N=N, d=d, metric="canonical", max_batch=B_FULL)
    f, g, h = fit.triple(xs, et)
    yield fit, f, g, h
    fit.close()


def test_c3_sample_vs_oracle(P, c3, c3_triple):
    T, xs, et = c3
    _, f, g, h = c3_triple
    spec = P.CostSpec(T, "full")
    for b in (0, 1, 95, 96, 383, 384, 511, 512, 1000, 1023):       # chunk boundaries of the host pipeline included
        fo, go, ho = P.triple(spec, list(xs[b]), list(et[b]), "canonical")
        assert abs(f[b] - fo) / fo < 1e-12
        assert rel(g[b], np.array(go)) < 1e-10
        assert rel(h[b], np.array(ho)) < 1e-10


def test_c3_tangency_and_finite(c3, c3_triple):
    _, xs, _ = c3
    _, f, g, h = c3_triple
    assert np.isfinite(f).all() and np.isfinite(g).all() and np.isfinite(h).all()
    assert (f > 0).all()
    assert np.abs(sym_part(xs, g)).max() < 1e-12 * max(1.0, np.abs(g).max())
    assert np.abs(sym_part(xs, h)).max() < 1e-12 * max(1.0, np.abs(h).max())


def test_c3_kraus_mixing_invariance(c3, c3_triple):
    T, xs, et = c3
    fit, f, g, h = c3_triple
    rng = np.random.default_rng(5)
    U = np.stack([random_orthogonal(16, rng) for _ in range(xs.shape[1])])
    f2, g2, h2 = fit.triple(mix_kraus(xs, U), mix_kraus(et, U))
    np.testing.assert_allclose(f2, f, rtol=1e-12)
    assert rel(g2, mix_kraus(g, U)) < 1e-10
    assert rel(h2, mix_kraus(h, U)) < 1e-10


def test_c3_hessian_symmetric_and_linear(c3, c3_triple):
    from opentn_b200 import stiefel
    _, xs, et = c3
    fit, _, g, h = c3_triple
    rng = np.random.default_rng(6)
    et2 = stiefel.project(xs, rng.standard_normal(xs.shape))
    fit.cost_grad(xs)
    h1 = fit.hvp_cached(et)
    h2 = fit.hvp_cached(et2)
    np.testing.assert_allclose(h1, h, rtol=0, atol=1e-12 * np.abs(h).max())
    a, b = cdot(xs, et2, h1), cdot(xs, et, h2)
    scale = np.sqrt(cdot(xs, h1, h1) * cdot(xs, et2, et2))
    assert (np.abs(a - b) / scale).max() < 1e-10
    h3 = fit.hvp_cached(2.0 * et - 3.0 * et2)
    assert rel(h3, 2.0 * h1 - 3.0 * h2) < 1e-11


def test_c3_first_directional_derivative_along_retraction(c3, c3_triple):
    """d/dt f(R_x(t eta)) at t = 0 is <grad f, eta> in the metric the gradient was taken in (central differences, all instances)."""
    from opentn_b200 import stiefel
    _, xs, et = c3
    fit, f, g, h = c3_triple
    t1 = 1e-5
    fp = fit.cost(stiefel.polar_decomposition_rectangular(xs, t1 * et))
    fm = fit.cost(stiefel.polar_decomposition_rectangular(xs, -t1 * et))
    d1 = (fp - fm) / (2 * t1)
    ge = cdot(xs, g, et)
    gn = np.sqrt(cdot(xs, g, g))                       # eta has canonical norm 1
    assert (np.abs(d1 - ge) / gn).max() < 1e-7


def test_c3_batch_independence_and_formulations(c3, c3_triple):
    """An instance gives the same numbers alone, in a small batch (dense chain chosen) or among 1024 (block-diagonal chain)."""
    from opentn_b200 import StiefelFit
    T, xs, et = c3
    _, f, g, h = c3_triple
    sl = slice(700, 732)
    for dense in (True, False):
        small = StiefelFit(T, model="full", N=N, d=d, metric="canonical", max_batch=32, dense=dense)
        fs, gs, hs = small.triple(xs[sl], et[sl])
        np.testing.assert_allclose(fs, f[sl], rtol=1e-12)
        assert rel(gs, g[sl]) < 1e-11
        assert rel(hs, h[sl]) < 1e-11
        f1, g1, h1 = small.triple(xs[777], et[777])
        assert abs(f1 - f[777]) / f[777] < 1e-12 and rel(g1, g[777]) < 1e-11 and rel(h1, h[777]) < 1e-11
        small.close()


def test_c3_dense_formulation_full_batch(c3, c3_triple):
    from opentn_b200 import StiefelFit
    T, xs, et = c3
    _, f, g, h = c3_triple
    dense = StiefelFit(T, model="full", N=N, d=d, metric="canonical", max_batch=B_FULL, dense=True)
    fd, gd, hd = dense.triple(xs, et)
    dense.close()
    np.testing.assert_allclose(fd, f, rtol=1e-12)
    assert rel(gd, g) < 1e-11
    assert rel(hd, h) < 1e-11


def test_c3_euclidean_metric_full_batch(P, c3):
    from opentn_b200 import StiefelFit, stiefel
    T, xs, et = c3
    fit = StiefelFit(T, model="full", N=N, d=d, metric="euclidean", max_batch=B_FULL)
    f, g, h = fit.triple(xs, et)
    assert np.abs(sym_part(xs, g)).max() < 1e-12 * max(1.0, np.abs(g).max())
    spec = P.CostSpec(T, "full")
    for b in (3, 640):
        fo, go, ho = P.triple(spec, list(xs[b]), list(et[b]), "euclidean")
        assert abs(f[b] - fo) / fo < 1e-12 and rel(g[b], np.array(go)) < 1e-10 and rel(h[b], np.array(ho)) < 1e-10
    edot = lambda a, b: np.einsum("blnp,blnp->b", a, b)
    # the Hessian is symmetric in the euclidean inner product
    et2 = stiefel.project(xs, np.random.default_rng(8).standard_normal(xs.shape))
    fit.cost_grad(xs)                                  # primal cache for hvp_cached (the chunked host path of triple keeps none)
    h2 = fit.hvp_cached(et2)
    assert (np.abs(edot(et2, h) - edot(et, h2)) / np.sqrt(edot(h, h) * edot(et2, et2))).max() < 1e-10
    # the polar retraction is a second-order retraction for the embedded (euclidean) metric (stiefel.py:323-355), so
    # d^2/dt^2 f(R_x(t eta)) at t = 0 is <eta, Hess f eta>_euclidean; all 1024 instances by central second differences
    t2 = 1e-3
    fp = fit.cost(stiefel.polar_decomposition_rectangular(xs, t2 * et))
    fm = fit.cost(stiefel.polar_decomposition_rectangular(xs, -t2 * et))
    d2 = (fp - 2 * f + fm) / t2 ** 2
    assert (np.abs(d2 - edot(et, h)) / np.sqrt(edot(h, h) * edot(et, et))).max() < 1e-4
    fit.close()


def test_c3_retraction_full_batch(c3):
    from opentn_b200 import stiefel
    _, xs, et = c3
    y = stiefel.polar_decomposition_rectangular(xs, 0.3 * et)
    gram = np.einsum("blnp,blnq->blpq", y, y)
    assert np.abs(gram - np.eye(16)).max() < 1e-13
    np.testing.assert_allclose(stiefel.polar_decomposition_rectangular(xs, np.zeros_like(xs)), xs, rtol=0, atol=1e-14)
    # polar factor is the closest isometry: (X+eta)^T Y is symmetric positive definite
    s = np.einsum("blnp,blnq->blpq", xs + 0.3 * et, y)
    assert np.abs(s - s.transpose(0, 1, 3, 2)).max() < 1e-12
    assert np.linalg.eigvalsh(s.reshape(-1, 16, 16)).min() > 0


def test_c3_tr_monotone_full_batch(c3):
    """Batched RTR on all 1024 instances: accepted steps never increase f, rejected ones keep x (trust_region_rcopt.py:76-85)."""
    from opentn_b200 import StiefelFit
    T, xs, _ = c3
    fit = StiefelFit(T, model="full", N=N, d=d, metric="canonical", max_batch=B_FULL)
    f0 = fit.cost(xs)
    x1, rep = fit.tr_run(xs, niter=4)
    f1 = fit.cost(x1)
    fit.close()
    assert (rep["status"] == 0).all()
    np.testing.assert_allclose(rep["f_iter"][0], f0, rtol=1e-12)
    assert (np.diff(rep["f_iter"], axis=0) <= 1e-15).all()
    assert (f1 <= rep["f_iter"][-1] + 1e-15).all()
    assert (f1 <= f0).all() and (f1 < f0).mean() > 0.99
    gram = np.einsum("blnp,blnq->blpq", x1, x1)
    assert np.abs(gram - np.eye(16)).max() < 1e-12


# ---------------------------------------------------------------------------------------------------------------------
# config-4 shape: 2-site model, m = 9, K = 16, 1024 instances
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def c4(P):
    Lk = P.create_2local_liouvillians(P.get_kitaev_nn_linbladian(1.0), N, d, pbc=True)[0]
    T = P.exp_operator_dt(Lk, 0.5)
    rng = np.random.default_rng(44)
    m, K = 9, 16
    g = rng.standard_normal((B_FULL, m, 4 * K, 4))
    q, r = np.linalg.qr(g)
    xs = q * np.sign(np.einsum("...ii->...i", r))[..., None, :]
    z = rng.standard_normal(xs.shape)
    s = np.einsum("blnp,blnq->blpq", xs, z)
    et = z - np.einsum("blnp,blpq->blnq", xs, 0.5 * (s + s.transpose(0, 1, 3, 2)))
    et /= np.sqrt(cdot(xs, et, et))[:, None, None, None]
    return T, xs, et


def test_c4_factored_vs_blocks_vs_oracle(P, c4):
    from opentn_b200 import StiefelFit
    T, xs, et = c4
    fact = StiefelFit(T, model="local", N=N, d=d, metric="canonical", max_batch=B_FULL, factored=True)
    f, g, h = fact.triple(xs, et)
    blocks = StiefelFit(T, model="local", N=N, d=d, metric="canonical", max_batch=B_FULL, dense=False)
    fb, gb, hb = blocks.triple(xs, et)
    blocks.close()
    np.testing.assert_allclose(fb, f, rtol=1e-12)
    assert rel(gb, g) < 1e-10
    assert rel(hb, h) < 1e-10
    spec = P.CostSpec(T, "local", N, d)
    for b in (0, 511, 1023):
        fo, go, ho = P.triple(spec, list(xs[b]), list(et[b]), "canonical")
        assert abs(f[b] - fo) / fo < 1e-12 and rel(g[b], np.array(go)) < 1e-10 and rel(h[b], np.array(ho)) < 1e-10
    assert np.abs(sym_part(xs, g)).max() < 1e-12 * max(1.0, np.abs(g).max())
    assert np.abs(sym_part(xs, h)).max() < 1e-12 * max(1.0, np.abs(h).max())
    # Kraus mixing
    rng = np.random.default_rng(9)
    U = np.stack([random_orthogonal(16, rng) for _ in range(xs.shape[1])])
    f2, g2, h2 = fact.triple(mix_kraus(xs, U), mix_kraus(et, U))
    np.testing.assert_allclose(f2, f, rtol=1e-11)
    assert rel(g2, mix_kraus(g, U)) < 1e-10
    assert rel(h2, mix_kraus(h, U)) < 1e-10
    # Hessian symmetry
    from opentn_b200 import stiefel
    et2 = stiefel.project(xs, rng.standard_normal(xs.shape))
    fact.cost_grad(xs)
    hh2 = fact.hvp_cached(et2)
    a, b_ = cdot(xs, et2, h), cdot(xs, et, hh2)
    assert (np.abs(a - b_) / np.sqrt(cdot(xs, h, h) * cdot(xs, et2, et2))).max() < 1e-10
    fact.close()
