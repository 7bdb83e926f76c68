"""Timing of the set-up stage that precedes the hot path: Liouvillian assembly and exp(tau L) on the device
against the host route (NumPy kron / SciPy expm) of the reference.  Usage: python tools/setup_profile.py [--n6-host]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import scipy.linalg

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opentn_b200 import transformations as T  # noqa: E402


def best(fn, reps=3):
    out = None
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        ts.append(time.perf_counter() - t0)
    return min(ts), out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n6-host", action="store_true", help="also time the host route at N=6 (tens of seconds)")
    args = ap.parse_args()
    res = {}
    Lnn = T.get_kitaev_nn_linbladian(1.0)
    for N in (4, 6):
        T.create_2local_liouvillians(Lnn, N, 2, pbc=True, backend="cuda")      # warm-up (context, staging buffers)
        t_dev, dev = best(lambda: T.create_2local_liouvillians(Lnn, N, 2, pbc=True, backend="cuda"))
        entry = {"liouvillian_device_s": t_dev}
        if N == 4 or args.n6_host:
            t_host, host = best(lambda: T.create_2local_liouvillians(Lnn, N, 2, pbc=True), reps=1 if N == 6 else 3)
            entry["liouvillian_host_s"] = t_host
            entry["max_abs_diff"] = float(max(np.abs(a - b).max() for a, b in zip(dev, host)))
        L = dev[0].real
        taus = [0.25, 0.5, 1.0, 2.0] if N == 4 else [4.0]
        T.exp_operator_sweep(L, taus)
        t_dev, e_dev = best(lambda: T.exp_operator_sweep(L, taus))
        entry["expm_sweep_taus"] = taus
        entry["expm_device_s"] = t_dev
        if N == 4 or args.n6_host:
            t_host, e_host = best(lambda: np.stack([scipy.linalg.expm(L * t) for t in taus]), reps=1 if N == 6 else 3)
            entry["expm_host_s"] = t_host
            entry["expm_rel_diff"] = float(np.abs(e_dev - e_host).max() / np.abs(e_host).max())
        res[f"N{N}"] = entry
    print(json.dumps(res))


if __name__ == "__main__":
    main()
