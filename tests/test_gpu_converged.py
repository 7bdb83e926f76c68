"""Converged-cost parity of the device trust-region solver (north star: "converged trust-region cost within 1e-8").

Two kinds of check:

* **Strict** -- on small, well-conditioned instances (target = model(x*) + noise, start = a kick away from x*) both the oracle's
  restatement of the reference loop (trust_region_rcopt.py:57-85 with the matrix-free Hessian operator) and `otn_tr_run` reach
  the same local minimiser with ||rgrad|| ~ 1e-8; their converged costs must agree to 1e-8 relative (observed ~1e-13).
* **Golden histories** -- the reference ships cost histories of long runs (`experiments/f_*.npy`).  Hessians at Trotter points
  are strongly indefinite and tCG exits flip on rounding-level margins, so two correct implementations separate
  after a handful of iterations and need not end in the same basin; what every correct run satisfies is: the recorded costs
  never increase (a rejected step keeps x), the first iterations coincide with the golden history, and the run ends at least
  as low as the reference's did up to the spread observed between equally valid trajectories (tolerances next to each case).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
N, d = 4, 2


@pytest.fixture(scope="module")
def P():
    from oracle import port
    return port


def _well_conditioned(P, m, K, seed, kick=0.05, noise=1e-3):
    rng = np.random.default_rng(seed)
    xstar = [P.random_stiefel(4 * K, 4, rng) for _ in range(m)]
    M = P.model_stiefel_local(xstar, N, d)
    T = M + noise * rng.standard_normal(M.shape)
    etas = [P.project(x, rng.standard_normal(x.shape)) for x in xstar]
    nrm = np.sqrt(P.canonical_dot(xstar, etas, etas))
    return T, P.retract_tangent(xstar, [kick * e / nrm for e in etas])


@pytest.mark.parametrize("m,K,seed", [(2, 2, 0), (3, 2, 0), (2, 3, 5)])
def test_converged_cost_matches_oracle(P, m, K, seed):
    from opentn_b200 import StiefelFit
    T, x0 = _well_conditioned(P, m, K, seed)
    spec = P.CostSpec(T, "local", N, d)
    niter = 25
    xs_o, f_o, _, _ = P.trust_region_fit(spec, x0, "canonical", niter=niter)
    g_o = np.linalg.norm(P.gradient_stiefel_vec(xs_o[-1], spec, "canonical"))
    fit = StiefelFit(T, model="local", N=N, d=d, metric="canonical")
    x_d, rep = fit.tr_run(np.array(x0), niter=niter)
    f_d, g_d = fit.cost_grad(x_d)
    gn_d = np.sqrt(P.canonical_dot(list(x_d), list(g_d), list(g_d)))
    assert g_o < 1e-6 and gn_d < 1e-6, (g_o, gn_d)                 # both runs are converged ...
    assert abs(f_d - spec(xs_o[-1])) / f_d < 1e-8                     # ... to the same cost (north star tolerance)
    # and the histories agree on the way there: to rounding level while every tCG exit is decided by a clear margin (the first
    # iterations), to 1e-6 afterwards (a tCG that stops one product earlier or later moves an intermediate cost by ~1e-8
    # relative -- observed 6e-8 -- without changing where the run converges)
    k = int(np.argmax(np.array(f_o) - f_o[-1] < 1e-9 * f_o[-1]))
    np.testing.assert_allclose(rep["f_iter"][:4, 0], f_o[:4], rtol=1e-9)
    np.testing.assert_allclose(rep["f_iter"][:k, 0], f_o[:k], rtol=1e-6)
    assert rep["status"][0] == 0
    fit.close()


def _run_segments(fit, xs0, segments):
    """Consecutive otn_tr_run calls warm-started with the radius of the previous one (the notebooks' `radius_init=` restarts)."""
    x, rad, fs, acc = np.array(xs0), None, [], []
    for nit in segments:
        x, rep = fit.tr_run(x, niter=nit, radius=rad)
        rad = rep["radius"]
        fs += list(rep["f_iter"][:, 0])
        acc += list(rep["accepted"][:, 0])
        assert rep["status"][0] == 0
    return x, np.array(fs), np.array(acc)


GOLDEN_RUNS = {
    # name: (builder of (T, xs0, model), golden key, segments, n_first: iterations that must coincide with the golden, end_tol)
    "c1_pspl_tau1_n1_K10": ("art/f_pauli_n1_tau1_rank10", [40, 20], 5),
    "c2_pspl_tau05_n4_K10": ("art/f_pspl_n4_t05_metric", [10, 5, 5], 3),
    "kitaev_tau05_n4_K16": ("art/f_kitaev_n4_rank_16", [30], 3),
    # (the stored full-sites history starts at the same cost but differs from iteration 1 on, 0.0706 vs 0.0710: its notebook settings
    #  -- radius, metric, Hessian basis -- are not recorded next to the .npy, so only the start and the end of the run are compared)
    "fullsites_kitaev_tau4_K4": ("art/data_trust_region_stiefel_r4_cost_fn", [50], 1),
}


def _problem(P, name):
    if name == "c1_pspl_tau1_n1_K10":
        Li = P.pspl_lindbladians(1.0)
        T = P.exp_operator_dt(P.create_2local_liouvillians(Li, N, d, pbc=True)[0], 1.0)
        return T, [P.super2ortho(x.real, rank=10) for x in P.get_general_trotter_local_ansatz(Li, 1, n=1)], "local"
    if name == "c2_pspl_tau05_n4_K10":
        Li = P.pspl_lindbladians(1.0)
        T = P.exp_operator_dt(P.create_2local_liouvillians(Li, N, d, pbc=True)[0], 0.5)
        return T, [P.super2ortho(x.real, rank=10) for x in P.get_general_trotter_local_ansatz(Li, 0.5, n=4)], "local"
    if name == "kitaev_tau05_n4_K16":
        Lk = [P.get_kitaev_nn_linbladian(1.0)]
        T = P.exp_operator_dt(P.create_2local_liouvillians(Lk[0], N, d, pbc=True)[0], 0.5)
        return T, [P.super2ortho(x.real, rank=16) for x in P.get_general_trotter_local_ansatz(Lk, 0.5, n=4)], "local"
    Lnn = P.get_kitaev_nn_linbladian(1.0)
    xs0 = [P.super2ortho(x.real, rank=4) for x in
           P.create_trotter_layers(list(P.create_2local_liouvillians(Lnn, N, d, pbc=False)), tau=4)[1:]]
    return P.kitaev_fullsites_target(), [xs0[0], xs0[1], xs0[0]], "full"


@pytest.mark.parametrize("name", list(GOLDEN_RUNS))
def test_golden_history_end_cost(P, goldens, name):
    """Run the device solver for as many iterations as the reference's stored history (same warm restarts)."""
    from opentn_b200 import StiefelFit
    key, segments, n_first = GOLDEN_RUNS[name]
    gold = np.asarray(goldens[key])
    T, xs0, model = _problem(P, name)
    fit = StiefelFit(T, model=model, N=N, d=d, metric="canonical")
    x, fs, acc = _run_segments(fit, xs0, segments)
    f_end = fit(x)
    # (1) the first iterations coincide with the reference's history (before rounding-level tCG decisions separate the runs)
    np.testing.assert_allclose(fs[:n_first], gold[:n_first], rtol=1e-9)
    # (2) recorded costs never increase: accepted steps have rho > rho_trust > 0 (a decrease), rejected steps keep x
    assert (np.diff(fs) <= 1e-13 * fs[:-1]).all()
    assert f_end <= fs[-1] * (1 + 1e-13)
    # (3) the run ends at least as low as the reference's own run of the same length, within the basin-to-basin spread
    #     (tolerance table in END_TOL: measured spread between the oracle's and the device's equally valid trajectories)
    k = min(len(fs), len(gold)) - 1
    assert fs[k] <= gold[k] * (1 + END_TOL[name]), (fs[k], gold[k])
    fit.close()


# relative slack of check (3).  Measured on a B200 with tools/converged_probe.py (profiles/converged_probe_r02.log), end of run,
# device vs reference history: C1 +13 % (0.005448 vs 0.004812; the ORACLE's own 60-iteration run, the reference loop restated on
# the CPU, ends at 0.005718 = +19 %: that is the spread between equally valid trajectories), C2 +4.4 %, Kitaev K=16 -0.9 %,
# full-sites K=4 -78 % (the device run keeps descending where the stored reference run stalls at 0.0413).
END_TOL = {
    "c1_pspl_tau1_n1_K10": 0.30,
    "c2_pspl_tau05_n4_K10": 0.15,
    "kitaev_tau05_n4_K16": 0.10,
    "fullsites_kitaev_tau4_K4": 0.0,
}
