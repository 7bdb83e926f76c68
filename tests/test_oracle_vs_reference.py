"""Oracle vs the UNMODIFIED reference executed under the jax->numpy shim.  Only runs where /root/reference exists (the build
container); on the GPU box the committed fixtures of tests/golden/ stand in for it."""
import io
import contextlib

import numpy as np
import pytest

from oracle.ref_shim import load_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not mounted")
N, d = 4, 2


@pytest.fixture(scope="module")
def ref():
    return load_reference()


def test_forward_path_bitwise(ref, oracle):
    P, T, O = oracle, ref.transformations, ref.optimization
    rng = np.random.default_rng(123)
    xs = [P.random_stiefel(20, 4, rng) for _ in range(7)]
    np.testing.assert_array_equal(np.asarray(O.model_stiefel_local(xs, N, d)), P.model_stiefel_local(xs, N, d))
    xf = [P.random_stiefel(48, 16, rng) for _ in range(3)]
    np.testing.assert_allclose(np.asarray(O.model_Ys_stiefel(xf)), P.model_Ys_stiefel(xf), atol=1e-14)
    Lnn = np.asarray(T.get_kitaev_nn_linbladian(0.7))
    np.testing.assert_array_equal(Lnn, P.get_kitaev_nn_linbladian(0.7))
    for pbc in (False, True):
        a = T.create_2local_liouvillians(Lnn, N, d, pbc)
        b = P.create_2local_liouvillians(Lnn, N, d, pbc)
        for u, v in zip(a, b):
            np.testing.assert_array_equal(np.asarray(u), v)
    assert O.compute_kitaev_approximation_error(d, N, 1.0, 4, 2) == pytest.approx(0.03642549, abs=5e-9)   # 2site cell 63


def test_reference_tr_loop_equals_oracle_loop(ref, oracle, goldens):
    """Same callables through the reference's loop and through the oracle's restatement of it: identical output."""
    P, S, TR = oracle, ref.stiefel, ref.trust_region_rcopt
    Tt = P.exp_operator_dt(P.create_2local_liouvillians(P.pspl_lindbladians(1.0), N, d, pbc=True)[0], 0.5)
    spec = P.CostSpec(Tt, "local", N, d)
    xs0 = [P.super2ortho(x.real, rank=10) for x in P.get_general_trotter_local_ansatz(P.pspl_lindbladians(1.0), 0.5, n=2)]
    gradf = lambda xi: P.gradient_stiefel_vec(xi, spec, "canonical")            # noqa: E731
    hessf = lambda xi: P.HessianOperator(spec, xi, "canonical")                  # noqa: E731
    with contextlib.redirect_stdout(io.StringIO()):
        xs_r, f_r, rad_r = TR.riemannian_trust_region_optimize(spec, S.retract_stiefel, gradf, hessf, xs0, save_x=True, niter=3)
    xs_o, f_o, rad_o, _ = P.riemannian_trust_region_optimize(spec, P.retract_stiefel, gradf, hessf, xs0, save_x=True, niter=3)
    np.testing.assert_array_equal(f_r, f_o)
    np.testing.assert_array_equal(np.array(xs_r), np.array(xs_o))
    assert rad_r == rad_o


def test_stiefel_algebra_bitwise(ref, oracle):
    P, S = oracle, ref.stiefel
    rng = np.random.default_rng(5)
    x = P.random_stiefel(24, 4, rng)
    z = rng.standard_normal((24, 4))
    e = S.project(x, rng.standard_normal((24, 4)))
    np.testing.assert_array_equal(S.project(x, z), P.project(x, z))
    for a0, a1 in ((1, 1), (1, 0.5)):
        np.testing.assert_array_equal(S.gradient_ambient2riemannian(x, z, a0, a1), P.gradient_ambient2riemannian(x, z, a0, a1))
        np.testing.assert_array_equal(S.riemannian_connection(z, e, e, x, a0, a1), P.riemannian_connection(z, e, e, x, a0, a1))
    v = np.asarray(S.parametrization_from_tangent(x, e))
    np.testing.assert_allclose(v, P.parametrization_from_tangent(x, e), atol=1e-14)
    np.testing.assert_array_equal(np.array(S.retract_stiefel([x], v)), np.array(P.retract_stiefel([x], v)))
