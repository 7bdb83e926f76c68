"""Generate `tests/golden/reference_goldens.npz` from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

Three kinds of fixtures are written:

(a) `art/<name>`     -- the reference's own stored artefacts (`/root/reference/experiments/*.npy`): optimised
                        isometries and cost histories.  Data files, copied verbatim (first iterates only for the
                        large history file).
(b) `known/<name>`   -- numbers printed in the reference's notebook outputs (cited cell by cell below).
(c) `ref/<name>`     -- outputs of the reference's own functions executed here under `oracle/ref_shim.py`
                        (jax -> numpy module shim) on seeded inputs that are stored alongside (`in/<name>`).

The reference tree does not exist on the GPU box; tests read only the .npz.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.ref_shim import REFERENCE_ROOT, load_reference  # noqa: E402
from oracle import port as P  # noqa: E402  (only for the derivative operator handed to the reference TR loop)

EXP = os.path.join(REFERENCE_ROOT, "experiments")


def main():
    ref = load_reference()
    T, O, S, TR = ref.transformations, ref.optimization, ref.stiefel, ref.trust_region_rcopt
    out = {}

    # ---------------------------------------------------------------- (a) stored artefacts
    art = [
        "xs_pspl_n1_rank_10", "f_pauli_n1_tau1_rank10", "f_pspl_n1_rank_10", "xs_timestep_1_opt_4", "f_timestep_1_opt_4",
        "xs_timestep_4_opt_9", "x_pauli_n2_rank_10_tau05", "x_pauli_n3_rank_10_tau05", "x_pauli_n4_rank_10_tau05",
        "f_pspl_n4_t05_metric", "x_kitaev_n4_rank_16", "f_kitaev_n4_rank_16", "x_kitaev_n4_rank_8",
        "xs_pspl_timestep_2_opt_4", "xs_pspl_timestep_3_opt_6", "xs_pspl_n1_rank_5", "xs_pspl_n1_rank_16",
        "data/trust_region_stiefel_r4_xs", "data/trust_region_stiefel_r4_cost_fn",
    ]
    for name in art:
        out["art/" + name.replace("data/", "data_")] = np.load(os.path.join(EXP, name + ".npy"))
    # first 8 iterates of the save_x=True history (trust_region_pauli_noise.ipynb cells 33, 143)
    out["art/xs_timestep_1_opt_first8"] = np.load(os.path.join(EXP, "xs_timestep_1_opt.npy"))[:8]
    out["art/xs_timestep_2_opt_first6"] = np.load(os.path.join(EXP, "xs_timestep_2_opt.npy"))[:6]

    # ---------------------------------------------------------------- (b) printed known answers
    known = {
        # trust_region_pauli_noise.ipynb cell 32 (cost at the Trotter point), cell 33 (first iterations, radii)
        "pspl_n1_tau1_K10_cost0": 0.11294517951068227,
        "pspl_n1_tau1_K10_first6": [0.11294517951068227, 0.11227828473846403, 0.11124027887142852,
                                    0.11014318479356361, 0.10950271927346032, 0.10944128179858796],
        # cell 112: Strang error vs n = 1..5 (PSPL N=4 tau=1)
        "pspl_strang_n1to5": [0.11294518, 0.02536439, 0.01085703, 0.00611177, 0.00393854],
        # cell 140/141
        "pspl_opt_cost_n1": 0.004330344959560802,
        "pspl_opt_cost_n4_print": 0.00013085,
        # trust_region_2site_ansatz.ipynb cell 63: Kitaev tau=4 Strang error n = 1..9
        "kitaev_strang_tau4_n1to9": [0.14312676, 0.03642549, 0.01624165, 0.00914625, 0.00585666, 0.00406828,
                                     0.00298945, 0.00228905, 0.00180877],
        # thesis_ranks.ipynb cells 15-20, 37: Kitaev N=4 tau=0.5 n=4 K=16
        "kitaev_n4_tau05_K16_cost0": 5.686635658263572e-05,
        # trust_region_fullsites_ansatz.ipynb cells 1, 6, 16: Kitaev N=4 tau=4 pbc=False, K=4
        "fullsites_cost0": 0.0959176702323516,
        # trust_region_higher_sites.ipynb cells 4, 16: N=6 Kitaev tau=4 n=1 K=2 pbc canonical
        "n6_cost0": 0.53944709,
        "n6_gradnorm": 15.82062161395626,
        # trust_region_taus.ipynb: Kitaev N=4, K=2, tau=0.5
        "kitaev_tau05_K2_n1": 0.0009094097438186602,
        "kitaev_tau05_K2_n2": 0.00022744282911377092,
        # trust_region_2site_ansatz.ipynb cell 113: parametrization_from_tangent known answer
        "param_from_tangent_cell113": [22.22289464, 3.67345833, 0, 0, 1.07103444, 0.22510018],
    }
    for k, v in known.items():
        out["known/" + k] = np.asarray(v, dtype=np.float64)

    # ---------------------------------------------------------------- (c) reference executed under the shim
    rng = np.random.default_rng(20241019)
    N, d = 4, 2

    # representation layer on random isometries (local K=3 and K=10; full-sites K=4)
    for tag, (n, p) in {"loc3": (12, 4), "loc10": (40, 4), "full4": (64, 16)}.items():
        x = P.random_stiefel(n, p, rng)
        out[f"in/x_{tag}"] = x
        out[f"ref/ortho2choi_{tag}"] = T.ortho2choi(x)
        out[f"ref/ortho2super_{tag}"] = T.ortho2super(x)
    K = 3
    out["ref/kraus2superop_loc3"] = T.kraus2superop([out["in/x_loc3"][k::K, :] for k in range(K)])

    # embedding (site B) incl. pbc shift, N=4 and N=6
    yl = rng.standard_normal((16, 16))
    out["in/yloc"] = yl
    for NN in (4, 6):
        st = T.create_supertensored_from_local(yl, NN, pbc=True)
        for shift in (False, True):
            full = T.convert_supertensored2liouvillianfull(st, NN, d, shift_pbc=shift)
            # store a strided sample for N=6 (the full 4096^2 would be 134 MB)
            out[f"ref/embed_N{NN}_shift{int(shift)}"] = full if NN == 4 else full[::61, ::67]

    # models
    xs_loc = [P.random_stiefel(12, 4, rng) for _ in range(5)]
    out["in/xs_loc5"] = np.array(xs_loc)
    out["ref/model_local_m5"] = O.model_stiefel_local(xs_loc, N, d)
    xs_full = [P.random_stiefel(32, 16, rng) for _ in range(3)]
    out["in/xs_full3"] = np.array(xs_full)
    out["ref/model_full_m3"] = np.asarray(O.model_Ys_stiefel(xs_full))
    out["ref/compose3"] = T.compose_superops_list([out["ref/ortho2super_full4"], out["ref/model_full_m3"],
                                                  out["ref/model_local_m5"]])

    # targets: PSPL tau in {1, 0.5}, Kitaev pbc tau=0.5, Kitaev obc tau=4 -- store norms + a row sample
    Li = P.pspl_lindbladians(1.0)
    Lnn = T.get_kitaev_nn_linbladian(1.0)
    out["ref/kitaev_Lnn"] = np.asarray(Lnn)
    targets = {}
    Lv = T.create_2local_liouvillians(Li, N, d, pbc=True)[0]
    targets["pspl_tau1"] = T.exp_operator_dt(Lv, 1.0)
    targets["pspl_tau05"] = T.exp_operator_dt(Lv, 0.5)
    Lk = T.create_kitaev_liouvillians(N, d, 1.0, pbc=True)[0]
    targets["kitaev_pbc_tau05"] = T.exp_operator_dt(Lk, 0.5)
    Lko = T.create_kitaev_liouvillians(N, d, 1.0, pbc=False)[0]
    targets["kitaev_obc_tau4"] = T.exp_operator_dt(Lko, 4.0)
    for k, v in targets.items():
        v = np.asarray(v)
        out[f"ref/target_{k}_fro"] = np.linalg.norm(v)
        out[f"ref/target_{k}_rows"] = v[::37]
        out[f"ref/target_{k}_imagmax"] = np.abs(v.imag).max()

    # Trotter starting points and costs at them
    xs0 = [T.super2ortho(x.real, rank=10) for x in O.get_general_trotter_local_ansatz(Li, 1, n=1)]
    out["ref/xs0_pspl_n1_K10"] = np.array(xs0)
    f_pspl = lambda xi: O.frobenius_norm(O.model_stiefel_local(xi, N, d), targets["pspl_tau1"])  # noqa: E731
    out["ref/cost_xs0_pspl_n1_K10"] = f_pspl(xs0)
    out["ref/strang_pspl_n1to5"] = np.array([O.compute_trotter_approximation_error(d, N, tau=1, n=n, Li=Li)
                                             for n in range(1, 6)])
    out["ref/strang_kitaev_tau4_n1to4"] = np.array([O.compute_kitaev_approximation_error(d, N, 1.0, 4, n)
                                                    for n in range(1, 5)])
    xs0k = [T.super2ortho(x.real, rank=16) for x in O.get_kitaev_trotter_local_ansatz(1.0, 0.5, 4)]
    out["ref/xs0_kitaev_tau05_n4_K16"] = np.array(xs0k)
    f_kit = lambda xi: O.frobenius_norm(O.model_stiefel_local(xi, N, d), targets["kitaev_pbc_tau05"])  # noqa: E731
    out["ref/cost_xs0_kitaev_tau05_n4_K16"] = f_kit(xs0k)
    # costs of the stored artefacts, evaluated by the reference's own cost code
    out["ref/cost_xs_pspl_n1_rank_10"] = f_pspl(list(out["art/xs_pspl_n1_rank_10"]))
    out["ref/cost_xs_timestep_1_opt_4"] = f_pspl(list(out["art/xs_timestep_1_opt_4"]))
    out["ref/cost_xs_timestep_4_opt_9"] = f_pspl(list(out["art/xs_timestep_4_opt_9"]))
    f_p05 = lambda xi: O.frobenius_norm(O.model_stiefel_local(xi, N, d), targets["pspl_tau05"])  # noqa: E731
    for n_ in (2, 3, 4):
        out[f"ref/cost_x_pauli_n{n_}_rank_10_tau05"] = f_p05(list(out[f"art/x_pauli_n{n_}_rank_10_tau05"]))
    out["ref/cost_x_kitaev_n4_rank_16"] = f_kit(list(out["art/x_kitaev_n4_rank_16"]))
    xr4 = [x.real for x in out["art/data_trust_region_stiefel_r4_xs"]]
    out["ref/cost_fullsites_r4_xs"] = O.frobenius_norm(O.model_Ys_stiefel(xr4), targets["kitaev_obc_tau4"])
    out["ref/cost_history_first8"] = np.array([f_pspl(list(x)) for x in out["art/xs_timestep_1_opt_first8"]])

    # Stiefel algebra on seeded inputs
    for tag, (n, p) in {"s40": (40, 4), "s64": (64, 16), "sq": (6, 6)}.items():
        x = P.random_stiefel(n, p, rng)
        z = rng.standard_normal((n, p))
        e = S.project(x, rng.standard_normal((n, p)))
        out[f"in/st_x_{tag}"], out[f"in/st_z_{tag}"], out[f"in/st_eta_{tag}"] = x, z, e
        out[f"ref/project_{tag}"] = S.project(x, z)
        out[f"ref/rgrad_euc_{tag}"] = S.gradient_ambient2riemannian(x, z, 1, 1)
        out[f"ref/rgrad_can_{tag}"] = S.gradient_ambient2riemannian(x, z, 1, 0.5)
        nu = S.gradient_ambient2riemannian(x, z, 1, 0.5)
        out[f"ref/conn_can_{tag}"] = S.riemannian_connection(z, nu, e, x, 1, 0.5)
        out[f"ref/conn_euc_{tag}"] = S.riemannian_connection(z, S.project(x, z), e, x, 1, 1)
        out[f"ref/canonical_metric_{tag}"] = S.canonical_metric(e, nu, x)
        out[f"ref/euclidean_metric_{tag}"] = S.euclidean_metric(e, nu)
        out[f"ref/polar_{tag}"] = S.polar_decomposition_rectangular(x, 0.3 * e)
        if n != p:
            xc = S.get_orthogonal_complement(x)
            out[f"ref/xcomp_{tag}"] = xc
            v = S.parametrization_from_tangent(x, e, x_comp=xc)
            out[f"ref/param_{tag}"] = np.asarray(v)
            A, B = S.parametrizations_from_vector(np.asarray(v), [x.shape])[0]
            out[f"ref/tangent_roundtrip_{tag}"] = x @ A + xc @ B
    # retraction of a 3-layer list through the coefficient vector (retract_stiefel)
    xs3 = [P.random_stiefel(40, 4, rng) for _ in range(3)]
    et3 = [0.05 * S.project(x, rng.standard_normal(x.shape)) for x in xs3]
    vec = np.hstack([np.asarray(S.parametrization_from_tangent(x, e)) for x, e in zip(xs3, et3)])
    out["in/retract_xs"], out["in/retract_etas"] = np.array(xs3), np.array(et3)
    out["ref/retract_vec"] = vec
    out["ref/retract_out"] = np.array(S.retract_stiefel(xs3, vec))

    # truncated CG on explicit matrices: SPD interior exit, SPD boundary exit, indefinite (negative curvature)
    for tag, (shift, radius) in {"interior": (30.0, 10.0), "boundary": (30.0, 0.05), "negcurv": (-5.0, 0.5)}.items():
        A = rng.standard_normal((30, 30))
        H = 0.5 * (A + A.T) + shift * np.eye(30)
        g = rng.standard_normal(30)
        z, onb = TR.truncated_cg(g, H, radius)
        out[f"in/tcg_H_{tag}"], out[f"in/tcg_g_{tag}"], out[f"in/tcg_radius_{tag}"] = H, g, radius
        out[f"ref/tcg_z_{tag}"], out[f"ref/tcg_onb_{tag}"] = z, onb

    # the reference's TR loop (unmodified) fed the oracle's matrix-free derivative operator: C1, 6 iterations
    spec = P.CostSpec(np.asarray(targets["pspl_tau1"]), "local", N, d)
    grad_vec = lambda xi: np.hstack([np.asarray(S.parametrization_from_tangent(x, g)) for x, g in  # noqa: E731
                                     zip(xi, P.rgrad(spec, xi, "canonical"))])
    hess_op = lambda xi: P.HessianOperator(spec, xi, "canonical")  # noqa: E731
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        xs_it, f_it, radius = TR.riemannian_trust_region_optimize(f_pspl, S.retract_stiefel, grad_vec, hess_op, xs0,
                                                                  save_x=True, niter=6)
    out["ref/tr_c1_f_iter"] = np.array(f_it)
    out["ref/tr_c1_x_iter"] = np.array(xs_it)
    out["ref/tr_c1_radius"] = radius

    path = os.path.join(HERE, "reference_goldens.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB,", len(out), "entries")
    print("TR C1 f_iter:", f_it)
    print("golden      :", out["art/f_pauli_n1_tau1_rank10"][:6])


if __name__ == "__main__":
    main()
