"""Pin the oracle's closed-form restatement of the reference's three autodiff call sites (stiefel.py:423,437/450 -> egrad;
stiefel.py:520 -> egrad_jvp) against (i) torch.func.jvp(torch.func.grad(.)) of a transcription of the reference cost and
(ii) central finite differences of the cost; plus the Riemannian layer against jvp of the rgrad map."""
import numpy as np
import pytest

from torch_autodiff import egrad_and_jvp

N, d = 4, 2


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / np.abs(np.asarray(b)).max()


@pytest.mark.parametrize("model,m,K", [("local", 3, 10), ("local", 5, 2), ("full", 3, 4), ("full", 1, 2)])
def test_closed_forms_vs_autodiff(oracle, model, m, K):
    P = oracle
    rng = np.random.default_rng(7)
    T = P.kitaev_fullsites_target() + (0.05j * rng.standard_normal((256, 256)) if model == "full" else 0)
    p = 16 if model == "full" else 4
    xs = [P.random_stiefel(p * K, p, rng) for _ in range(m)]
    et = [P.project(x, rng.standard_normal(x.shape)) for x in xs]
    spec = P.CostSpec(T, model, N, d)
    f, Z, Zd = P.egrad_jvp(spec, xs, et)
    f2, g2, gd2 = egrad_and_jvp(T, model, xs, et, N, d)
    assert abs(f - f2) / f2 < 1e-13
    for a, b in zip(Z, g2):
        assert rel(a, b) < 1e-12
    for a, b in zip(Zd, gd2):
        assert rel(a, b) < 1e-12
    f3, Z3 = P.cost_and_egrad(spec, xs)
    assert f3 == f and all(np.array_equal(a, b) for a, b in zip(Z, Z3))


def test_egrad_vs_finite_differences(oracle):
    P = oracle
    rng = np.random.default_rng(3)
    T = P.exp_operator_dt(P.create_2local_liouvillians(P.pspl_lindbladians(1.0), N, d, pbc=True)[0], 1.0)
    xs = [P.random_stiefel(12, 4, rng) for _ in range(3)]
    spec = P.CostSpec(T, "local", N, d)
    Z = P.egrad(spec, xs)
    h = 1e-6
    for (l, i, j) in [(0, 3, 1), (1, 7, 2), (2, 11, 3)]:
        xp = [x.copy() for x in xs]; xm = [x.copy() for x in xs]
        xp[l][i, j] += h; xm[l][i, j] -= h
        fd = (spec(xp) - spec(xm)) / (2 * h)
        assert abs(fd - Z[l][i, j]) < 1e-8                                 # trust_region_2site_ansatz.ipynb cells 120-122


@pytest.mark.parametrize("metric", ["euclidean", "canonical"])
def test_riemannian_layer(oracle, metric):
    """rgrad special cases == general formula (trust_region_2site_ansatz.ipynb cells 15-16); HVP is symmetric in the
    coefficient basis at a critical-point-free generic x only up to the connection term, but its matrix must equal the dense
    column-by-column construction (stiefel.py:498-544)."""
    P = oracle
    rng = np.random.default_rng(11)
    T = P.kitaev_fullsites_target()
    xs = [P.random_stiefel(32, 16, rng) for _ in range(2)]
    spec = P.CostSpec(T, "full")
    g = P.rgrad(spec, xs, metric)
    Z = P.egrad(spec, xs)
    for x, z, gg in zip(xs, Z, g):
        want = P.project(x, z) if metric == "euclidean" else z - x @ z.T @ x     # stiefel.py:438, 451
        np.testing.assert_allclose(gg, want, atol=1e-13)
        assert np.allclose(x.T @ gg, -(x.T @ gg).T, atol=1e-12)                   # tangent
    op = P.HessianOperator(spec, xs, metric)
    v = rng.standard_normal(sum(P.stiefel_tangent_dof(*x.shape) for x in xs))
    w = rng.standard_normal(v.shape)
    # linearity of the operator
    np.testing.assert_allclose(op @ (2 * v + w), 2 * (op @ v) + (op @ w), atol=1e-10)
    # the Riemannian Hessian is self-adjoint w.r.t. the metric that defines the gradient; for the canonical metric that is
    # exactly the coefficient dot product
    if metric == "canonical":
        assert abs(w @ (op @ v) - v @ (op @ w)) < 1e-9 * abs(w @ (op @ v))
