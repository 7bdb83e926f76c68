"""GPU parity of the complex-isometry path (OTN_FLAG_COMPLEX, csrc/complex.cuh) against oracle/complex_port.py, which is itself
pinned against torch autodiff (tests/test_oracle_complex.py).  Tolerances as for the real path: cost 1e-12, gradient and
Hessian-vector product 1e-10 relative as (n,p) tangent matrices."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def CP():
    from oracle import complex_port
    return complex_port


def _problem(CP, N, m, K, B, seed, noise=0.05):
    dim = 2 ** N
    rng = np.random.default_rng(seed)
    xs = np.array([[CP.random_stiefel(dim * K, dim, rng) for _ in range(m)] for _ in range(B)])
    et = np.array([[CP.project(x, rng.standard_normal(x.shape) + 1j * rng.standard_normal(x.shape)) for x in xb] for xb in xs])
    D = dim * dim
    T = CP.model([CP.random_stiefel(dim * K, dim, rng) for _ in range(m)]) + noise * (rng.standard_normal((D, D)) + 1j * rng.standard_normal((D, D)))
    return T, xs, et


@pytest.mark.parametrize("metric", ["canonical", "euclidean", (2.0, 0.7)])
@pytest.mark.parametrize("N,m,K", [(4, 3, 2), (4, 2, 4), (2, 3, 3), (2, 1, 2)])
def test_complex_triple_vs_oracle(CP, metric, N, m, K):
    from opentn_b200 import StiefelFit
    B = 2
    T, xs, et = _problem(CP, N, m, K, B, seed=10 * N + m + K)
    fit = StiefelFit(T, model="full", N=N, d=2, metric=metric, max_batch=B)
    f, g, h = fit.triple(xs, et)
    assert g.dtype == np.complex128 and h.dtype == np.complex128 and f.dtype == np.float64
    for b in range(B):
        fo, go, ho = CP.triple(T, list(xs[b]), list(et[b]), metric)
        assert abs(f[b] - fo) / fo < 1e-12
        assert rel(g[b], np.array(go)) < 1e-10
        assert rel(h[b], np.array(ho)) < 1e-10
    # cost only, cost + gradient (+ Euclidean gradient in the convention df = Re tr(Z^H dx)), cached HVP
    np.testing.assert_allclose(fit.cost(xs), f, rtol=1e-13)
    f2, g2, e2 = fit.cost_grad(xs, want_egrad=True)
    np.testing.assert_allclose(f2, f, rtol=1e-13)
    assert rel(g2, g) < 1e-13
    assert rel(e2[0], np.array(CP.egrad_jvp(T, list(xs[0]))[1])) < 1e-10
    assert rel(fit.hvp_cached(et), h) < 1e-12
    # single-instance call forms (list of isometries in, float + arrays out)
    f1, g1 = fit.cost_grad(list(xs[1]))
    assert isinstance(f1, float) and abs(f1 - f[1]) / f1 < 1e-13 and rel(g1, g[1]) < 1e-13
    fit.close()


def test_complex_reduces_to_real_path(CP):
    """Real isometries with a COMPLEX target: the complex handle (forced by an input with a rounding-level imaginary part) gives the
    real path's numbers -- Im T only adds a constant under the root in the real path, while the complex path sees it as data."""
    from opentn_b200 import StiefelFit
    from oracle import port as P
    rng = np.random.default_rng(4)
    xs = np.array([[P.random_stiefel(32, 16, rng) for _ in range(3)]])
    et = np.array([[P.project(x, rng.standard_normal(x.shape)) for x in xs[0]]])
    T = P.kitaev_fullsites_target()
    fit = StiefelFit(T, model="full", N=4, d=2, metric="canonical")
    fr, gr, hr = fit.triple(xs, et)
    xc = xs.astype(complex); xc[0, 0, 0, 0] += 1e-300j            # complex dtype with a non-zero imaginary entry -> complex handle
    fc, gc, hc = fit.triple(xc, et.astype(complex) + 0j * 1e-300)
    assert gc.dtype == np.complex128
    assert abs(fr[0] - fc[0]) / fr[0] < 1e-13
    assert np.abs(gc.imag).max() < 1e-13 and rel(gc.real, gr) < 1e-12
    assert np.abs(hc.imag).max() < 1e-12 and rel(hc.real, hr) < 1e-11
    fit.close()


def test_complex_retraction(CP):
    import scipy.linalg
    from opentn_b200 import complex_stiefel as CS
    rng = np.random.default_rng(7)
    X = np.array([CP.random_stiefel(24, 4, rng) for _ in range(5)])
    Z = np.array([0.2 * CP.project(x, rng.standard_normal(x.shape) + 1j * rng.standard_normal(x.shape)) for x in X])
    out = CS.polar_decomposition_rectangular(X, Z)
    for b in range(5):
        np.testing.assert_allclose(out[b], scipy.linalg.polar(X[b] + Z[b])[0], atol=1e-13)
        np.testing.assert_allclose(out[b].conj().T @ out[b], np.eye(4), atol=1e-13)
    np.testing.assert_allclose(CS.polar_decomposition_rectangular(X[:1], 5.0 * X[:1])[0], X[0], atol=1e-13)   # Gram = 36: scaled into the basin
    with pytest.raises(np.linalg.LinAlgError):
        CS.polar_decomposition_rectangular(X[:1], -X[:1])              # X + Z = 0: singular Gram


def test_complex_coefficient_vectors(CP):
    from opentn_b200 import complex_stiefel as CS
    rng = np.random.default_rng(8)
    x = CP.random_stiefel(12, 4, rng)
    a = CP.project(x, rng.standard_normal(x.shape) + 1j * rng.standard_normal(x.shape))
    b = CP.project(x, rng.standard_normal(x.shape) + 1j * rng.standard_normal(x.shape))
    va, vb = CS.parametrization_from_tangent(x, a), CS.parametrization_from_tangent(x, b)
    assert va.shape == (CS.tangent_dof(12, 4),)
    assert abs(va @ vb - CP.canonical_dot([x], [a], [b])) < 1e-13
    np.testing.assert_allclose(CS.tangent_from_parametrization(x, va), a, atol=1e-13)
    np.testing.assert_array_equal(va, CP.parametrization_from_tangent(x, a))          # same basis convention as the oracle


def test_complex_trust_region_converges_to_oracle_cost(CP):
    """Host trust-region loop (reference signature) over the GPU callables for complex isometries vs the same loop over the CPU
    oracle's callables: same history while the steps are well determined, converged cost within 1e-8 (north star tolerance)."""
    from opentn_b200 import StiefelFit
    from opentn_b200 import complex_stiefel as CS
    from opentn_b200.trust_region_rcopt import riemannian_trust_region_optimize
    N, m, K, dim = 2, 2, 2, 4
    rng = np.random.default_rng(11)
    xstar = [CP.random_stiefel(dim * K, dim, rng) for _ in range(m)]
    T = CP.model(xstar) + 1e-3 * (rng.standard_normal((16, 16)) + 1j * rng.standard_normal((16, 16)))
    kick = [CP.project(x, rng.standard_normal(x.shape) + 1j * rng.standard_normal(x.shape)) for x in xstar]
    nrm = np.sqrt(CP.canonical_dot(xstar, kick, kick))
    x0 = CP.retract(xstar, [0.05 * e / nrm for e in kick])
    fit = StiefelFit(T, model="full", N=N, d=2, metric="canonical")
    f, retract, gradfunc, hessfunc = CS.closures(fit)
    xs_d, fs_d, _ = riemannian_trust_region_optimize(f, retract, gradfunc, hessfunc, x0, niter=20, verbose=False)

    class OracleHess:
        def __init__(self, xs):
            self.xs = xs
            self.comps = [CP.get_orthogonal_complement(x) for x in xs]

        def __matmul__(self, v):
            sizes = [CS.tangent_dof(*x.shape) for x in self.xs]
            etas, o = [], 0
            for x, c, sz in zip(self.xs, self.comps, sizes):
                etas.append(CP.tangent_from_parametrization(x, v[o:o + sz], c)); o += sz
            h = CP.triple(T, self.xs, etas, "canonical")[2]
            return np.concatenate([CP.parametrization_from_tangent(x, z, c) for x, z, c in zip(self.xs, h, self.comps)])

    def o_grad(xs):
        g = CP.cost_rgrad(T, xs, "canonical")[1]
        return np.concatenate([CP.parametrization_from_tangent(x, z) for x, z in zip(xs, g)])

    def o_retract(xs, eta):
        sizes = [CS.tangent_dof(*x.shape) for x in xs]
        etas, o = [], 0
        for x, sz in zip(xs, sizes):
            etas.append(CP.tangent_from_parametrization(x, eta[o:o + sz])); o += sz
        return CP.retract(xs, etas)
    xs_o, fs_o, _ = riemannian_trust_region_optimize(lambda xs: CP.cost(T, xs), o_retract, o_grad, OracleHess, x0, niter=20, verbose=False)
    np.testing.assert_allclose(fs_d[:4], fs_o[:4], rtol=1e-9)
    f_d, f_o = fit.cost(np.stack(xs_d[-1])), CP.cost(T, xs_o[-1])
    g_end = np.linalg.norm(gradfunc(xs_d[-1]))
    assert g_end < 1e-6, g_end
    assert abs(f_d - f_o) / f_o < 1e-8
    assert f_d < fs_d[0]
    fit.close()
