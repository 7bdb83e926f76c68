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


# ---- tau sweeps (otn_expm_batch) and device Liouvillian assembly (otn_liouvillian) ---------------------------------------

def test_expm_sweep_matches_scipy(oracle):
    from opentn_b200 import transformations as T
    P = oracle
    Lfull, Lodd, Leven = P.create_2local_liouvillians(P.get_kitaev_nn_linbladian(1.0), 4, 2, pbc=True)
    taus = [0.25, 0.5, 1.0, 2.0]                      # the tau grid of BASELINE config 4
    sweep = T.create_trotter_layers_sweep([Lfull, Lodd, Leven], taus)
    for k, tau in enumerate(taus):
        want = [scipy.linalg.expm(Lfull * tau), scipy.linalg.expm(Lodd * tau / 2), scipy.linalg.expm(Leven * tau)]
        for got, w in zip(sweep[k], want):
            assert got.shape == w.shape and got.dtype == w.dtype
            assert relmax(got, w) < 1e-12
    # one tau through the batch entry point == the single-matrix entry point (same scaling power)
    one = T.exp_operator_sweep(Lfull, [2.0])[0]
    np.testing.assert_array_equal(one, T.exp_operator_dt(Lfull, 2.0, backend="cuda"))


def test_expm_sweep_complex_and_negative_tau():
    from opentn_b200 import transformations as T
    rng = np.random.default_rng(5)
    A = (rng.standard_normal((48, 48)) + 1j * rng.standard_normal((48, 48))) / 7
    taus = [-1.0, 0.0, 0.3, 3.0]
    got = T.exp_operator_sweep(A, taus)
    for g, t in zip(got, taus):
        assert relmax(g, scipy.linalg.expm(A * t)) < 1e-11
    with pytest.raises(Exception):
        T.exp_operator_sweep(A, [np.nan])


@pytest.mark.parametrize("N,pbc", [(2, True), (4, False), (4, True)])
@pytest.mark.parametrize("model", ["kitaev", "pspl", "random_complex"])
def test_liouvillian_assembly(oracle, N, pbc, model):
    """Device assembly == the reference's kron route (oracle restatement, bitwise-pinned to the reference in
    tests/test_oracle_golden.py); products of two doubles summed in a different order: 1e-15 absolute."""
    from opentn_b200 import transformations as T
    P = oracle
    if model == "kitaev":
        Li = P.get_kitaev_nn_linbladian(0.7)
    elif model == "pspl":
        Li = P.pspl_lindbladians(1.3)
    else:
        rng = np.random.default_rng(N)
        Li = [rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4)) for _ in range(3)]
    want = P.create_2local_liouvillians(Li, N, 2, pbc=pbc)
    got = T.create_2local_liouvillians(Li, N, 2, pbc=pbc, backend="cuda")
    for g, w in zip(got, want):
        assert g.shape == w.shape
        assert np.abs(g - w).max() <= 4e-15 * max(1.0, np.abs(w).max())


def test_liouvillian_qutrits_and_limits():
    """d = 3 against the host kron route of this package; argument checks of the C entry point."""
    from opentn_b200 import transformations as T
    from opentn_b200._lib import OtnError
    rng = np.random.default_rng(0)
    L = rng.standard_normal((9, 9)) + 1j * rng.standard_normal((9, 9))
    want = T.vectorize_dissipative(np.kron(L, np.eye(3))) + T.vectorize_dissipative(np.kron(np.eye(3), L))
    got = T._liouvillian_terms_device([L], [(0, 0, 1), (0, 1, 2)], 3, 3)
    assert np.abs(got - want).max() < 1e-13
    with pytest.raises(OtnError):
        T._liouvillian_terms_device([L], [(0, 0, 0)], 3, 3)      # a pair needs two different sites
    with pytest.raises(OtnError):
        T._liouvillian_terms_device([L], [(1, 0, 1)], 3, 3)      # operator index out of range


def test_liouvillian_n6_trace_preserving():
    """N = 6 (D = 4096): too large for the host route in a test; check the defining properties instead: <<1|L = 0
    (trace preservation), L is real for the real Kitaev jump operator, and odd/even parts commute site-wise structure
    L = odd + even."""
    from opentn_b200 import transformations as T
    full, odd, even, _ = T.create_kitaev_liouvillians(6, 2, 1.0, pbc=True, backend="cuda")
    vec_id = np.eye(64).reshape(-1)
    assert np.abs(vec_id @ full).max() < 1e-13
    assert np.abs(full.imag).max() == 0.0
    # every term acts on two sites: the odd part of N = 6 restricted to sites (0,1) equals the N = 2 dissipator
    two = T.create_kitaev_liouvillians(2, 2, 1.0, backend="numpy")[0]
    # partial "trace" check: apply odd to |rho_01 (x) 1/16>> and compare with two applied to rho_01 per pair
    rng = np.random.default_rng(1)
    r = rng.standard_normal((4, 4)); r = r @ r.T
    rho = np.kron(r, np.eye(16)) / 16
    out = (odd @ rho.reshape(-1)).reshape(64, 64)
    # pairs (2,3) and (4,5) see the identity, on which the Kitaev dissipator acts as D[1] (not zero in general)
    dI = (two @ np.eye(4).reshape(-1)).reshape(4, 4)
    want = (np.kron((two @ r.reshape(-1)).reshape(4, 4), np.eye(16)) + np.kron(r, np.kron(dI, np.eye(4)))
            + np.kron(r, np.kron(np.eye(4), dI))) / 16
    assert np.abs(out - want).max() < 1e-13
