"""GPU parity of the batched device-side trust region (otn_tr_run) and of the host-loop mirror
(opentn_b200.trust_region_rcopt + stiefel) against the oracle's restatement of trust_region_rcopt.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
N, d = 4, 2


@pytest.fixture(scope="module")
def P():
    from oracle import port
    return port


@pytest.fixture(scope="module")
def c1(P):
    Li = P.pspl_lindbladians(1.0)
    T = P.exp_operator_dt(P.create_2local_liouvillians(Li, N, d, pbc=True)[0], 1.0)
    xs0 = [P.super2ortho(x.real, rank=10) for x in P.get_general_trotter_local_ansatz(Li, 1, n=1)]
    return T, xs0


def test_tr_c1_trajectory(P, c1, goldens):
    """C1 (PSPL N=4 tau=1 n=1 K=10, canonical metric, defaults): the device solver reproduces the reference's golden
    cost trajectory for the iterations over which tCG exit decisions coincide (0-4), and the oracle's per-step log."""
    from opentn_b200 import StiefelFit
    T, xs0 = c1
    fit = StiefelFit(T, model="local", N=N, d=d, metric="canonical", dense=False)    # block-diagonal path; the host-loop test below takes the dense one
    x, rep = fit.tr_run(np.array(xs0), niter=5, save_x=True)
    np.testing.assert_allclose(rep["f_iter"][:, 0], goldens["art/f_pauli_n1_tau1_rank10"][:5], rtol=1e-10)
    xs_it, f_it, radius, log_ = P.trust_region_fit(P.CostSpec(T, "local", N, d), xs0, "canonical", niter=5, save_x=True)
    np.testing.assert_allclose(rep["f_iter"][:, 0], f_it, rtol=1e-11)
    # steps 0-3 agree in every decision; in step 4 the tCG exit is decided by a rounding-level margin ("a
    # 1e-16 perturbation flipped a tCG exit at TR-iteration 5"), so from there only the cost *before* the step is compared
    k = 4
    np.testing.assert_allclose(rep["rho"][:k, 0], [l["rho"] for l in log_[:k]], rtol=1e-6)
    np.testing.assert_array_equal(rep["radius_iter"][:k, 0], [l["radius"] for l in log_[:k]])
    np.testing.assert_array_equal(rep["on_boundary"][:k, 0], [int(l["on_boundary"]) for l in log_[:k]])
    np.testing.assert_array_equal(rep["accepted"][:k, 0], [int(l["accepted"]) for l in log_[:k]])
    # n_matvec of the oracle operator counts tCG products + the model-value product (which the device solver accumulates
    # from the tCG products instead of recomputing)
    np.testing.assert_array_equal(rep["n_cg"][:k, 0] + 1, [l["n_matvec"] for l in log_[:k]])
    assert rep["hvp_total"] == rep["n_cg"].sum()
    np.testing.assert_allclose(rep["x_iter"][k, 0], np.array(xs_it[k]), atol=1e-8)
    assert rep["status"][0] == 0
    assert x.shape == (3, 40, 4) and np.allclose(x, rep["x_iter"][-1, 0])
    fit.close()


def test_tr_single_steps_from_stored_iterates(c1, goldens):
    """One device TR step from the reference's stored iterate k lands on its stored iterate k+1."""
    from opentn_b200 import StiefelFit
    T, _ = c1
    fit = StiefelFit(T, model="local", N=N, d=d, metric="canonical", max_batch=4)
    hist = goldens["art/xs_timestep_1_opt_first8"]
    radii = np.array([0.01, 0.02, 0.04, 0.08])
    x, rep = fit.tr_run(hist[:4], niter=1, radius=radii)     # the four steps as one batch of four instances
    np.testing.assert_allclose(x, hist[1:5], atol=1e-9)
    np.testing.assert_array_equal(rep["radius"], [0.02, 0.04, 0.08, 0.1])
    fit.close()


@pytest.mark.parametrize("dense", [False, True])
def test_tr_batch_matches_oracle_fullsites(P, dense):
    """Full-sites ansatz, random starts (headline workload shape, K=4 to keep the oracle fast): per-step parity for a batch."""
    from opentn_b200 import StiefelFit
    T = P.kitaev_fullsites_target()
    B, K = 3, 4
    rng = np.random.default_rng(5)
    xs = np.array([[P.random_stiefel(16 * K, 16, rng) for _ in range(3)] for _ in range(B)])
    fit = StiefelFit(T, model="full", N=N, d=d, metric="canonical", max_batch=B, dense=dense)
    x, rep = fit.tr_run(xs, niter=2)
    spec = P.CostSpec(T, "full")
    for b in range(B):
        xs_it, f_it, radius, log_ = P.trust_region_fit(spec, list(xs[b]), "canonical", niter=2)
        np.testing.assert_allclose(rep["f_iter"][:, b], f_it, rtol=1e-10)
        np.testing.assert_array_equal(rep["accepted"][:, b], [int(l["accepted"]) for l in log_])
        np.testing.assert_allclose(x[b], np.array(xs_it[-1]), atol=1e-8)
    fit.close()


@pytest.mark.parametrize("dense", [False, True])
def test_tr_rejected_steps_in_a_batch(P, dense):
    """Oversized trust region: two of the three instances reject their second step, the third accepts every step.  From the
    second iteration on the solver reuses the forward pass of f(x_next) for accepted instances and redoes it for rejected ones
    (otn_dev_cost_grad_cached_forward); costs, decisions and iterates must follow the oracle through and after the rejection."""
    from opentn_b200 import StiefelFit
    T = P.kitaev_fullsites_target()
    B, K, niter = 3, 2, 4
    rng = np.random.default_rng(6)
    xs = np.array([[P.random_stiefel(16 * K, 16, rng) for _ in range(3)] for _ in range(B)])
    fit = StiefelFit(T, model="full", N=N, d=d, metric="canonical", max_batch=B, dense=dense)
    x, rep = fit.tr_run(xs, niter=niter, radius_init=3.0, maxradius=6.0)
    spec = P.CostSpec(T, "full")
    acc = []
    for b in range(B):
        xs_it, f_it, radius, log_ = P.trust_region_fit(spec, list(xs[b]), "canonical", niter=niter, radius_init=3.0, maxradius=6.0)
        np.testing.assert_allclose(rep["f_iter"][:, b], f_it, rtol=1e-10)
        np.testing.assert_array_equal(rep["accepted"][:, b], [int(l["accepted"]) for l in log_])
        np.testing.assert_allclose(rep["rho"][:, b], [l["rho"] for l in log_], rtol=1e-6)
        np.testing.assert_allclose(x[b], np.array(xs_it[-1]), atol=1e-8)
        acc.append([int(l["accepted"]) for l in log_])
    acc = np.array(acc)
    assert (acc == 0).any() and (acc[:, 1:] == 1).all(axis=1).any(), "the case must mix rejected and accepted steps"
    fit.close()


def test_host_loop_mirror(P, c1, goldens):
    """The reference-signature host loop (riemannian_trust_region_optimize with gradient_stiefel_vec /
    riemannian_hessian_vec / retract_stiefel closures) over the GPU callables."""
    from opentn_b200 import StiefelFit
    from opentn_b200.stiefel import gradient_stiefel_vec, retract_stiefel, riemannian_hessian_vec
    from opentn_b200.trust_region_rcopt import riemannian_trust_region_optimize
    T, xs0 = c1
    f_stiefel = StiefelFit(T, model="local", N=N, d=d, metric="canonical")
    grad_stiefel = lambda xi: gradient_stiefel_vec(xi, f_stiefel, metric="canonical")        # noqa: E731
    hessian_stiefel = lambda xi: riemannian_hessian_vec(xi, f_stiefel, metric="canonical")   # noqa: E731
    xs, fs, radius = riemannian_trust_region_optimize(f_stiefel, retract_stiefel, grad_stiefel, hessian_stiefel, xs0,
                                                      save_x=True, niter=4, verbose=False)
    np.testing.assert_allclose(fs, goldens["art/f_pauli_n1_tau1_rank10"][:4], rtol=1e-10)
    assert radius == 0.1 or radius == 0.08 or radius > 0
    assert len(xs) == 5
    f_stiefel.close()


def test_sweep_driver_small(P):
    """Reduced config-4 sweep (2 taus x 2 ranks x 3 seeds, 2 TR iterations): runs through the sharding-aware driver and every
    instance's first cost equals the oracle's cost at its kicked starting point."""
    from opentn_b200 import sweep
    from opentn_b200 import transformations as T
    Lnn = [T.get_kitaev_nn_linbladian(1.0)]
    taus, ranks, seeds = [0.5, 1.0], [2, 4], [0, 1, 2]
    res = sweep.run_sweep(Lnn, N, d, taus, ranks, seeds, n_steps=1, niter=2)
    Lvec = P.create_2local_liouvillians(P.get_kitaev_nn_linbladian(1.0), N, d, pbc=True)[0]
    for K in ranks:
        r = res[K]
        assert len(r["f0"]) == 6 and (r["status"] == 0).all()
        assert (r["f_final"] <= r["f0"] * (1 + 1e-12)).all()
        for i in (0, 5):
            tau = taus[r["tau_index"][i]]
            # starting points come from the device factory (otn_super2ortho: Kraus operators up to sign) + a 1e-2 kick drawn by
            # the device generator (otn_kicked_starts); the oracle restates both (Philox4x32-10 in NumPy)
            x0 = r["x0"][i]
            layers_i = np.stack([np.asarray(x).real for x in P.get_general_trotter_local_ansatz([P.get_kitaev_nn_linbladian(1.0)], tau, 1)])
            want = P.kicked_start(list(T.super2ortho_batch(layers_i, K)), sweep.instance_seed(K, r["tau_index"][i], r["seed"][i]))
            np.testing.assert_allclose(x0, np.array(want), rtol=0, atol=1e-12)
            np.testing.assert_allclose(np.swapaxes(x0, -1, -2) @ x0, np.broadcast_to(np.eye(4), (x0.shape[0], 4, 4)), atol=1e-13)
            f = P.CostSpec(P.exp_operator_dt(Lvec, tau), "local", N, d)(list(x0))
            assert abs(f - r["f0"][i]) / f < 1e-10
            # the un-kicked device start reproduces the host-built Trotter start's cost (same superoperators, Kraus signs aside)
            layers = np.stack([np.asarray(x).real for x in P.get_general_trotter_local_ansatz([P.get_kitaev_nn_linbladian(1.0)], tau, 1)])
            xs_dev = T.super2ortho_batch(layers, K)
            xs_host = [P.super2ortho(x, rank=K) for x in layers]
            spec = P.CostSpec(P.exp_operator_dt(Lvec, tau), "local", N, d)
            assert abs(spec(list(xs_dev)) - spec(xs_host)) < 1e-9


def test_kicked_starts_batched_equals_single(P):
    from opentn_b200 import sweep
    rng = np.random.default_rng(0)
    xs0 = np.stack([np.linalg.qr(rng.standard_normal((16, 4)))[0] for _ in range(5)])
    many = sweep.kicked_starts(np.stack([xs0, xs0, xs0]), [3, 8, 3])
    one = sweep.kicked_start(list(xs0), 8)
    np.testing.assert_allclose(many[1], one, rtol=0, atol=1e-14)
    np.testing.assert_array_equal(many[0], many[2])
