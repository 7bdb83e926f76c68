"""otn_expm (SURVEY 8f "next" #1) against scipy.linalg.expm, the routine the reference's exp_operator_dt calls
(transformations.py:24-32).  Floating-point kernel: tolerance 1e-12 relative to ||exp||_max."""
import numpy as np
import pytest
import scipy.linalg

pytestmark = pytest.mark.gpu


def relmax(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


@pytest.mark.parametrize("tau", [0.25, 1.0, 4.0])
@pytest.mark.parametrize("model", ["kitaev_pbc", "kitaev_obc", "pspl"])
def test_liouvillian_targets(oracle, tau, model):
    from opentn_b200 import transformations as T
    P = oracle
    if model == "pspl":
        L = P.create_2local_liouvillians(P.pspl_lindbladians(1.0), 4, 2, pbc=True)[0]
    else:
        L = P.create_2local_liouvillians(P.get_kitaev_nn_linbladian(1.0), 4, 2, pbc=(model == "kitaev_pbc"))[0]
    want = scipy.linalg.expm(L * tau)
    got = T.exp_operator_dt(L, tau, backend="cuda")
    assert got.dtype == L.dtype
    assert relmax(got, want) < 1e-12
    # and the fit sees the same cost with either target
    np.testing.assert_allclose(T.exp_operator_dt(L, tau, backend="scipy"), want)


@pytest.mark.parametrize("D", [16, 64, 100, 144])
def test_random_complex_and_odd_sizes(D):
    from opentn_b200 import transformations as T
    rng = np.random.default_rng(D)
    A = rng.standard_normal((D, D)) / np.sqrt(D)
    for M in (A, A + 1j * rng.standard_normal((D, D)) / np.sqrt(D), 30 * A):
        want = scipy.linalg.expm(M * 0.7)
        got = T.exp_operator_dt(M, 0.7, backend="cuda")
        assert relmax(got, want) < 1e-11


def test_semigroup_and_trace_preservation(oracle):
    """Size-independent properties: exp((s+t)L) = exp(sL) exp(tL); a Liouvillian propagator preserves the trace
    (<<1| exp(tau L) = <<1| in row-major vectorisation)."""
    from opentn_b200 import transformations as T
    P = oracle
    L = P.create_2local_liouvillians(P.get_kitaev_nn_linbladian(1.0), 4, 2, pbc=True)[0].real
    a, b, c = (T.exp_operator_dt(L, t, backend="cuda") for t in (0.5, 1.5, 2.0))
    assert relmax(a @ b, c) < 1e-12
    vec_id = np.eye(16).reshape(-1)
    np.testing.assert_allclose(vec_id @ c, vec_id, atol=1e-12)
