"""Prints device TR cost histories next to the reference's golden histories (sets the tolerances of tests/test_gpu_converged.py)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from opentn_b200 import StiefelFit
from oracle import port as P

g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "reference_goldens.npz"))
N, d = 4, 2
np.set_printoptions(precision=10, linewidth=200)

def run(name, T, xs0, model, gold, segments):
    fit = StiefelFit(T, model=model, N=N, d=d, metric="canonical")
    x = np.array(xs0); rad = None; fs = []
    t = time.time()
    for nit in segments:
        x, rep = fit.tr_run(x, niter=nit, radius=rad)
        rad = rep["radius"]
        fs += list(rep["f_iter"][:, 0])
    dt = time.time() - t
    ffin = fit(x)
    print(f"== {name}: {sum(segments)} its in {dt:.2f}s, final radius {rad}, f(x_final) {ffin:.12g}")
    k = min(len(fs), len(gold))
    print(" device:", np.array(fs)[:k])
    print(" golden:", np.asarray(gold)[:k])
    print(" n_cg last segment:", rep["n_cg"][:, 0], "accepted:", rep["accepted"][:, 0])
    fit.close()

Li = P.pspl_lindbladians(1.0)
T1 = P.exp_operator_dt(P.create_2local_liouvillians(Li, N, d, pbc=True)[0], 1.0)
xs0 = [P.super2ortho(x.real, rank=10) for x in P.get_general_trotter_local_ansatz(Li, 1, n=1)]
run("C1 pspl tau1 n1 K10", T1, xs0, "local", g["art/f_pauli_n1_tau1_rank10"], [40, 20])
T05 = P.exp_operator_dt(P.create_2local_liouvillians(Li, N, d, pbc=True)[0], 0.5)
xs0 = [P.super2ortho(x.real, rank=10) for x in P.get_general_trotter_local_ansatz(Li, 0.5, n=4)]
run("C2 pspl tau05 n4 K10", T05, xs0, "local", g["art/f_pspl_n4_t05_metric"], [10, 5, 5])
Lk = [P.get_kitaev_nn_linbladian(1.0)]
Tk = P.exp_operator_dt(P.create_2local_liouvillians(Lk[0], N, d, pbc=True)[0], 0.5)
xs0 = [P.super2ortho(x.real, rank=16) for x in P.get_general_trotter_local_ansatz(Lk, 0.5, n=4)]
run("kitaev tau05 n4 K16", Tk, xs0, "local", g["art/f_kitaev_n4_rank_16"], [30])
Lnn = P.get_kitaev_nn_linbladian(1.0)
Tf = P.kitaev_fullsites_target()
xs0f = [P.super2ortho(x.real, rank=4) for x in P.create_trotter_layers(list(P.create_2local_liouvillians(Lnn, N, d, pbc=False)), tau=4)[1:]]
xs0f = [xs0f[0], xs0f[1], xs0f[0]]
run("fullsites kitaev tau4 K4", Tf, xs0f, "full", g["art/data_trust_region_stiefel_r4_cost_fn"], [50])
