"""CPU tests of the host-side mirror (opentn_b200.transformations / optimization / stiefel / trust_region_rcopt): problem
set-up against the reference's golden vectors, coefficient-basis glue, and the host TR/tCG control flow (fed CPU callables
from the oracle, so no GPU is needed)."""
import numpy as np
import pytest

N, d = 4, 2


def test_setup_layer_matches_goldens(goldens):
    from opentn_b200 import transformations as T
    np.testing.assert_array_equal(T.get_kitaev_nn_linbladian(1.0), goldens["ref/kitaev_Lnn"])
    Lp = T.create_2local_liouvillians([np.kron(T.X, T.I) - np.kron(T.I, T.X), np.kron(T.Z, T.I) - np.kron(T.I, T.Z)], N, d, pbc=True)[0]
    for tau, name in ((1.0, "pspl_tau1"), (0.5, "pspl_tau05")):
        t = T.exp_operator_dt(Lp, tau)
        assert np.isclose(np.linalg.norm(t), goldens[f"ref/target_{name}_fro"], rtol=1e-14)
        np.testing.assert_allclose(t[::37], goldens[f"ref/target_{name}_rows"], atol=1e-14)
    Lk, Lo, Le, Lnn = T.create_kitaev_liouvillians(N, d, 1.0, pbc=True)
    np.testing.assert_allclose(Lk, Lo + Le)                                       # tests/test_transformations.py:85-92
    t = T.exp_operator_dt(Lk, 0.5)
    np.testing.assert_allclose(t[::37], goldens["ref/target_kitaev_pbc_tau05_rows"], atol=1e-14)
    t = T.exp_operator_dt(T.create_kitaev_liouvillians(N, d, 1.0, pbc=False)[0], 4.0)
    np.testing.assert_allclose(t[::37], goldens["ref/target_kitaev_obc_tau4_rows"], atol=1e-14)
    with pytest.raises(ValueError, match="not a valid backend"):
        T.exp_operator_dt(Lk, 1.0, backend="torch")


def test_starting_points_match_goldens(goldens):
    from opentn_b200 import optimization as O
    from opentn_b200 import transformations as T
    Li = [np.kron(T.X, T.I) - np.kron(T.I, T.X), np.kron(T.Z, T.I) - np.kron(T.I, T.Z)]
    xs0 = [T.super2ortho(x.real, rank=10) for x in O.get_general_trotter_local_ansatz(Li, 1, n=1)]
    np.testing.assert_allclose(np.array(xs0), goldens["ref/xs0_pspl_n1_K10"], atol=1e-14)
    xs0k = [T.super2ortho(x.real, rank=16) for x in O.get_kitaev_trotter_local_ansatz(1.0, 0.5, 4)]
    np.testing.assert_allclose(np.array(xs0k), goldens["ref/xs0_kitaev_tau05_n4_K16"], atol=1e-13)
    assert len(xs0k) == 9 and xs0k[0].shape == (64, 4)


@pytest.mark.parametrize("dd,NN", [(2, 4), (2, 6), (3, 3)])
def test_choi_super_involution(dd, NN):
    """tests/test_transformations.py:16-20"""
    from opentn_b200 import transformations as T
    size = dd ** (2 * NN) if NN < 6 else 256
    dim = int(np.sqrt(size))
    C = np.random.default_rng(0).standard_normal((size, size))
    np.testing.assert_array_equal(T.super2choi(T.choi2super(C, dim=dim), dim=dim), C)


def test_ortho_choi_roundtrip_and_psd(goldens):
    from opentn_b200 import transformations as T
    x = goldens["in/x_loc10"]
    np.testing.assert_array_equal(T.ortho2choi(x), goldens["ref/ortho2choi_loc10"])
    np.testing.assert_array_equal(T.choi2ortho(T.ortho2choi(x)), x)
    A = np.random.default_rng(1).standard_normal((16, 16))
    Cm = A @ A.T
    Xf = T.factorize_psd(Cm)
    np.testing.assert_allclose(Xf @ Xf.T, Cm, atol=1e-12)                          # tests/test_transformations.py:48-54
    with pytest.raises(ValueError, match="invalid eigenvalue found"):              # tests/test_transformations.py:56-62
        T.factorize_psd(-np.eye(16))
    np.testing.assert_allclose(T.unfactorize_psd(T.factorize_psd_truncated(Cm, chi_max=16)), Cm, atol=1e-12)


@pytest.mark.parametrize("dd,NN", [(2, 4), (3, 4)])
def test_permutation_roundtrip(dd, NN):
    """tests/test_transformations.py:71-74"""
    from opentn_b200 import transformations as T
    size = dd ** (2 * NN)
    S = np.random.default_rng(2).standard_normal((size, size)) if size <= 256 else None
    if S is None:
        pytest.skip("6561^2 is too large for the CPU suite")
    np.testing.assert_array_equal(T.convert_liouvillianfull2supertensored(T.convert_supertensored2liouvillianfull(S, NN, dd), NN, dd), S)


def test_embedding_host_matches_goldens(goldens):
    from opentn_b200 import transformations as T
    st = T.create_supertensored_from_local(goldens["in/yloc"], 4, pbc=True)
    for shift in (False, True):
        np.testing.assert_array_equal(T.convert_supertensored2liouvillianfull(st, 4, 2, shift_pbc=shift),
                                      goldens[f"ref/embed_N4_shift{int(shift)}"])
    with pytest.raises(AssertionError, match="Only even number of sites"):        # transformations.py:414
        T.create_supertensored_from_local(goldens["in/yloc"], 3)


def test_stiefel_index_helpers():
    """tests/test_stiefel.py of the reference."""
    from opentn_b200.stiefel import get_ij_unit_matrix, get_k_unit_matrix, int2tuple, tuple2int
    with pytest.raises(AssertionError, match="out of range"):
        tuple2int(0, 3, 2)
    for i, j, p in [(1, 4, 5), (3, 2, 4)]:
        assert int2tuple(tuple2int(i, j, p), p) == (i, j)
    with pytest.raises(AssertionError, match="out of range"):
        get_k_unit_matrix(3, 3, 10)
    for d0, d1, k in [(3, 4, 9), (4, 3, 2), (5, 5, 10)]:
        i, j = int2tuple(k, cols=d1)
        assert get_k_unit_matrix(d0, d1, k)[i, j] == 1
        np.testing.assert_array_equal(get_k_unit_matrix(d0, d1, k), get_ij_unit_matrix(d0, d1, i, j))


def test_coefficient_glue_roundtrip(oracle, goldens):
    """(A-upper, B) parametrisation with our own X_perp: round trip, dof counts, and dot == canonical metric."""
    from opentn_b200 import stiefel as S
    x, e = goldens["in/st_x_s40"], goldens["in/st_eta_s40"]
    xc = S.get_orthogonal_complement(x)
    assert xc.shape == (40, 36)
    np.testing.assert_allclose(xc.T @ xc, np.eye(36), atol=1e-13)
    np.testing.assert_allclose(x.T @ xc, 0, atol=1e-13)
    v = S.parametrization_from_tangent(x, e)
    assert v.shape == (S.stiefel_tangent_dof(x),) == (150,)
    A, B = S.parametrizations_from_vector(v, [x.shape])[0]
    np.testing.assert_allclose(S.tangent_from_parametrization(x, A, B), e, atol=1e-13)
    assert abs(v @ v - oracle.canonical_dot([x], [e], [e])) < 1e-12 * (v @ v)
    # known answer of trust_region_2site_ansatz.ipynb cell 113 has the same structure: [triu(A), B.ravel()]
    assert goldens["known/param_from_tangent_cell113"].shape == (6,)
    xs = goldens["in/st_x_sq"]
    assert np.all(S.get_orthogonal_complement(xs) == 0)
    assert S.parametrization_from_tangent(xs, xs @ S.antisymmetrize(np.arange(36.).reshape(6, 6))).shape == (15,)


@pytest.mark.parametrize("tag", ["interior", "boundary", "negcurv"])
def test_truncated_cg_host(goldens, tag):
    from opentn_b200.trust_region_rcopt import truncated_cg
    z, onb = truncated_cg(goldens[f"in/tcg_g_{tag}"], goldens[f"in/tcg_H_{tag}"], float(goldens[f"in/tcg_radius_{tag}"]))
    np.testing.assert_array_equal(z, goldens[f"ref/tcg_z_{tag}"])
    assert onb == bool(goldens[f"ref/tcg_onb_{tag}"])


def test_quadratic_solver_errors():
    from opentn_b200.trust_region_rcopt import solve_quadratic_equation
    with pytest.raises(ValueError, match="non-negative discriminant"):             # trust_region_rcopt.py:159-160
        solve_quadratic_equation(1.0, 2.0)
    assert solve_quadratic_equation(0.0, -4.0) == (-2.0, 2.0)
    lo, hi = solve_quadratic_equation(1.5, -4.0)
    assert lo < hi and abs(hi * hi + 3 * hi - 4) < 1e-12


def test_host_tr_loop_with_cpu_callables(oracle, goldens):
    """The host TR loop (same control flow as trust_region_rcopt.py:5-89) reproduces the reference's loop output when fed
    the oracle's CPU callables -- isolates the control flow from the GPU arithmetic."""
    from opentn_b200.trust_region_rcopt import riemannian_trust_region_optimize
    P = oracle
    Tt = P.exp_operator_dt(P.create_2local_liouvillians(P.pspl_lindbladians(1.0), N, d, pbc=True)[0], 1.0)
    spec = P.CostSpec(Tt, "local", N, d)
    xs0 = list(goldens["ref/xs0_pspl_n1_K10"])
    xs, fs, radius = riemannian_trust_region_optimize(
        spec, P.retract_stiefel, lambda xi: P.gradient_stiefel_vec(xi, spec, "canonical"),
        lambda xi: P.HessianOperator(spec, xi, "canonical"), xs0, save_x=True, niter=4, verbose=False)
    np.testing.assert_allclose(fs, goldens["ref/tr_c1_f_iter"][:4], rtol=1e-12)
    np.testing.assert_allclose(np.array(xs), goldens["ref/tr_c1_x_iter"][:5], atol=1e-10)
    assert len(xs) == 5 and radius == 0.1
    # save_x=False keeps only the first and the last iterate (trust_region_rcopt.py:80-81)
    xs2, fs2, _ = riemannian_trust_region_optimize(
        spec, P.retract_stiefel, lambda xi: P.gradient_stiefel_vec(xi, spec, "canonical"),
        lambda xi: P.HessianOperator(spec, xi, "canonical"), xs0, niter=2, verbose=False)
    assert len(xs2) == 2 and len(fs2) == 2


def test_gradient_requires_fit_object():
    from opentn_b200.stiefel import gradient_stiefel_vec, riemannian_hessian_vec
    with pytest.raises(TypeError, match="StiefelFit"):
        gradient_stiefel_vec([np.zeros((8, 4))], lambda x: 0.0)
    with pytest.raises(ValueError, match="not a valid metric value"):
        riemannian_hessian_vec([np.zeros((8, 4))], None, metric="bogus")


def test_elementary_tangent_directions_and_predicates():
    """Host mirrors of the reference's tangent-basis helpers (stiefel.py:112-173, 252-264, 376-396): the k-th elementary direction
    is the tangent whose coefficient vector is the k-th unit vector -- the convention the dense Hessian's columns follow."""
    from opentn_b200 import stiefel as S
    rng = np.random.default_rng(0)
    x = np.linalg.qr(rng.standard_normal((8, 4)))[0]
    dof = S.stiefel_tangent_dof(x)
    assert dof == 6 + 16
    for k in range(dof):
        z = S.get_elementary_tangent_direction(k, x)
        assert S.is_in_tangent_space(x, z)
        e = np.zeros(dof); e[k] = 1.0
        np.testing.assert_allclose(S.parametrization_from_tangent(x, z), e, atol=1e-12)
    with pytest.raises(AssertionError):
        S.get_elementary_tangent_direction(dof, x)
    q = np.linalg.qr(rng.standard_normal((4, 4)))[0]
    a = S.parametrization_from_tangent_square(q, q @ np.array([[0, 1, 0, 0], [-1, 0, 0, 2], [0, 0, 0, 0], [0, -2, 0, 0.]]))
    np.testing.assert_allclose(a, [1, 0, 0, 0, 2, 0], atol=1e-12)
    assert S.is_hermitian(np.eye(3)) is True
    ok, dist = S.is_hermitian(np.array([[0, 1], [0, 0.]]))
    assert ok is False and dist == pytest.approx(np.sqrt(2))
    assert S.is_symmetric(np.eye(2)) and S.is_antisymmetric(np.array([[0, 1], [-1, 0.]])) and not S.is_antisymmetric(np.eye(2))
    assert [u.shape for u in S.get_unit_matrices([x, q])] == [(8, 4), (4, 4)] and S.get_unit_matrices([x])[0][0, 0] == 1
    p = S.random_psd(5)
    assert p.trace() == pytest.approx(1.0) and np.linalg.eigvalsh(p).min() > -1e-12


def test_liouvillian_term_lists_match_kron_route(monkeypatch):
    """The (operator, site0, site1) term lists `create_2local_liouvillians(backend='cuda')` hands to otn_liouvillian, evaluated
    here by a NumPy stand-in for the kernel, reproduce the reference's kron / permute route (odd, even and periodic sums)."""
    from opentn_b200 import transformations as T

    def place(op, s0, s1, N, d):
        # two-site operator with its first factor on site s0 and its second on s1, identity elsewhere
        full = np.kron(op, np.eye(d ** (N - 2)))                       # factors on sites (0, 1), rest in order
        rest = [s for s in range(N) if s not in (s0, s1)]
        perm = [0] * N                                                  # new axis j <- old axis perm[j]
        for old, new in enumerate([s0, s1] + rest):
            perm[new] = old
        return T.permute_operator(full, perm, d)

    def numpy_terms(ops, terms, N, d):
        out = np.zeros((d ** (2 * N), d ** (2 * N)), dtype=complex)
        for j, s0, s1 in terms:
            out += T.vectorize_dissipative(place(np.asarray(ops[j], dtype=complex), s0, s1, N, d))
        return out

    monkeypatch.setattr(T, "_liouvillian_terms_device", numpy_terms)
    rng = np.random.default_rng(0)
    Li = [rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4)), T.get_kitaev_nn_linbladian(0.5)]
    for N, pbc in ((2, True), (4, False), (4, True)):
        want = T.create_2local_liouvillians(Li, N, 2, pbc=pbc)
        got = T.create_2local_liouvillians(Li, N, 2, pbc=pbc, backend="cuda")
        for g, w in zip(got, want):
            np.testing.assert_allclose(g, w, rtol=0, atol=1e-13)
    with pytest.raises(ValueError):
        T.create_2local_liouvillians(Li, 3, 2, backend="cuda")


def test_row_shard_argument_checks():
    """StiefelFit.row_shard (BASELINE config 5) validates on the host before anything touches a GPU: it shards the dense chain
    only, needs an initialised process group, and a world size that divides D into row blocks of whole 64-row tiles."""
    import torch.distributed as dist
    from opentn_b200 import StiefelFit
    T = np.eye(256)
    with pytest.raises(ValueError, match="dense=True"):
        StiefelFit(T, model="local", N=4, d=2).row_shard()
    with pytest.raises(ValueError, match="dense=True"):
        StiefelFit(T, model="local", N=4, d=2, factored=True).row_shard()
    assert not dist.is_initialized()
    with pytest.raises(RuntimeError, match="process group"):
        StiefelFit(T, model="local", N=4, d=2, dense=True).row_shard()
