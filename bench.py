#!/usr/bin/env python
"""Benchmark of the Stiefel trust-region hot path on B200 (contract: see the round brief / DESIGN.md section "Measurement").

Workload (BASELINE.json configs[2], SURVEY 8d "C3"): Kitaev chain N=4 (D=256), full-sites ansatz, m=3 layers, Kraus rank 16,
x [B,3,256,16], B = 1024 random-init instances per GPU, target exp(4 L) (pbc=False).  One *step* = one unit of work U1 per
instance of the batch: (f, rgrad, H eta) = cost + Riemannian gradient + Hessian-vector product sharing the primal pass.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm
  python bench.py --impl reference [...]                           # the reference's CPU algorithm (oracle port) on host cores

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cost+grad+HVP evals/sec (n=4 rank-16 Kitaev full-sites batch)"
UNIT = "evals/s"
N_SITES, D_LOC, M_LAYERS, K_RANK, BATCH = 4, 2, 3, 16, 1024
DIM = D_LOC ** N_SITES
NROWS, D_SUP = DIM * K_RANK, DIM * DIM
# algorithmic FLOPs per unit U1: 9(m-1) real D^3 products + assembly and its reverse/tangent variants
FLOPS_PER_TRIPLE = 9 * (M_LAYERS - 1) * 2 * D_SUP ** 3 + 6 * M_LAYERS * 2 * K_RANK * D_SUP ** 2
GEMM_FLOPS_PER_TRIPLE = 9 * (M_LAYERS - 1) * 2 * D_SUP ** 3


def workload_config(extra=None):
    cfg = {"workload": "C3: Kitaev gamma=1 N=4 tau=4 obc, full-sites ansatz m=3 K=16, x[B,3,256,16], B=1024 random-init "
                       "instances per GPU, unit U1 = (f, rgrad, H eta)",
           "batch_per_gpu": BATCH, "N": N_SITES, "D": D_SUP, "m": M_LAYERS, "K": K_RANK, "metric_tensor": "canonical",
           "cache": "inputs + work matrices per step = 10.9 GB >> 126 MB L2 (no flush needed)"}
    if extra:
        cfg.update(extra)
    return cfg


def make_inputs(first, count):
    """Instances first..first+count-1 of C3: x = qf(G) with G ~ N(0,1), default_rng(1234+b), sign-fixed QR;
    eta = P_x(G'), default_rng(10^6+b), scaled to canonical norm 1.  Pure NumPy host set-up (not timed)."""
    import numpy as np
    xs = np.empty((count, M_LAYERS, NROWS, DIM))
    et = np.empty_like(xs)
    for i in range(count):
        rng = np.random.default_rng(1234 + first + i)
        for l in range(M_LAYERS):
            q, r = np.linalg.qr(rng.standard_normal((NROWS, DIM)))
            xs[i, l] = q * np.sign(np.diag(r))
        rng2 = np.random.default_rng(10 ** 6 + first + i)
        nrm = 0.0
        for l in range(M_LAYERS):
            x = xs[i, l]
            g = rng2.standard_normal((NROWS, DIM))
            xtg = x.T @ g
            e = g - x @ (0.5 * (xtg + xtg.T))
            et[i, l] = e
            xe = x.T @ e
            nrm += np.sum(e * e) - 0.5 * np.sum(xe * xe)
        et[i] /= np.sqrt(nrm)
    return xs, et


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle port, NumPy) on the host cores
# ---------------------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    first, count = args
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    from oracle import port as P
    T = P.kitaev_fullsites_target(N_SITES, D_LOC, 1.0, 4.0, False)
    spec = P.CostSpec(T, "full")
    xs, et = make_inputs(first, count)
    P.triple(spec, list(xs[0]), list(et[0]), "canonical")  # warm the BLAS / page in
    t0 = time.perf_counter()
    acc = 0.0
    for i in range(count):
        f, g, h = P.triple(spec, list(xs[i]), list(et[i]), "canonical")
        acc += f
    return time.perf_counter() - t0, acc


def _cpu_tr_worker(args):
    """One trust-region iteration (matrix-free Hessian operator, BASELINE.md baseline (ii)) per instance on the CPU oracle."""
    first, count = args
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    import numpy as np
    from oracle import port as P
    T = P.kitaev_fullsites_target(N_SITES, D_LOC, 1.0, 4.0, False)
    spec = P.CostSpec(T, "full")
    xs, _ = make_inputs(first, count)
    t0 = time.perf_counter()
    for i in range(count):
        P.trust_region_fit(spec, list(xs[i]), "canonical", niter=1)
    t_tr = time.perf_counter() - t0
    # baseline (i): the reference builds the dense Hessian one column per tangent degree of freedom (stiefel.py:498-529);
    # time a few columns of exactly that construction and let the caller extrapolate to all of them
    op = P.HessianOperator(spec, list(xs[0]), "canonical")
    size = sum(P.stiefel_tangent_dof(*x.shape) for x in xs[0])
    ncol = 4
    t1 = time.perf_counter()
    for k in range(ncol):
        e = np.zeros(size); e[k * (size // ncol)] = 1.0
        op @ e
    return t_tr, (time.perf_counter() - t1) / ncol * size


class CpuPool:
    """Worker pool for the oracle port: one instance per process, min(B, cores) processes, 1 BLAS thread each
    (BASELINE.md section 3 protocol).  Spawned once; every `run` is one bounded sample."""

    def __init__(self, cores=None):
        import multiprocessing as mp
        self.cores = cores or os.cpu_count() or 1
        self.pool = mp.get_context("spawn").Pool(self.cores)
        self.pool.map(_cpu_worker, [(0, 1)] * self.cores)  # spawn + import + BLAS warm-up
        self.offset = 0

    def run(self, sample_per_core):
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_worker, [(self.offset + c * sample_per_core, sample_per_core) for c in range(self.cores)])
        wall = time.perf_counter() - t0
        self.offset = (self.offset + self.cores * sample_per_core) % 4096
        inner = max(r[0] for r in res)
        total = self.cores * sample_per_core
        return total / inner, total, wall

    def run_tr(self, sample_per_core):
        res = self.pool.map(_cpu_tr_worker, [(c * sample_per_core, sample_per_core) for c in range(self.cores)])
        dense_iter_s = statistics.median(r[1] for r in res)      # seconds per TR iteration spent building the dense Hessian
        return self.cores * sample_per_core / max(r[0] for r in res), self.cores / dense_iter_s

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_triples_per_sec(sample_per_core=16, cores=None):
    pool = CpuPool(cores)
    try:
        v, total, wall = pool.run(sample_per_core)
    finally:
        pool.close()
    return v, pool.cores, total, wall


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    per_core = 16
    pool = CpuPool()
    cores = pool.cores
    try:
        for _ in range(args.warmup):
            pool.run(2)
        values, t_ms = [], []
        for _ in range(args.steps):
            v, total, wall = pool.run(per_core)
            values.append(v)
            t_ms.append(1e3 * total / v)
    finally:
        pool.close()
    v = statistics.median(values)
    sample = (f"{cores * per_core} instances of C3 per step ({per_core} per process x {cores} processes, 1 BLAS thread each); "
              f"value = median over {args.steps} steps")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": statistics.median(t_ms), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference = opentn's algorithm restated in NumPy (oracle/port.py; JAX is not installable here), "
                    "matrix-free HVP (BASELINE.md baseline (ii)); CPU work does not scale with --gpus"}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, smax, reasons, power = [], [], set(), []
        for ln in out.strip().splitlines():
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); smax.append(float(parts[2])); power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "power_w_max": max(power), "samples": len(sm)}


def measure_fp64_peak(torch):
    """cuBLAS DGEMM 8192^3 through torch.matmul -- the measuring stick for the FP64 tensor roofline (MEASURED_PEAKS.json
    carries bf16 only).  Library call used as a ruler, never on the product path."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    torch.cuda.empty_cache()
    return 2 * n ** 3 / best / 1e9  # TFLOP/s


def bind_to_gpu_numa_node(gpu_index):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (sysfs), so that the pinned host buffers of the end-to-end
    leg are first-touched there; with 8 ranks streaming 400 MB per step each, remote-node traffic is what limits `e2e`.
    Returns a description for the JSON line, or None when the topology is not exposed."""
    if os.environ.get("OTN_BENCH_NO_NUMA"):
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(gpu_index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dev = "/sys/bus/pci/devices/" + bus.lower()[-12:]
        node = int(open(dev + "/numa_node").read())
        if node < 0:
            return None
        cpulist = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        cpus = set()
        for part in cpulist.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return {"node": node, "cpus": cpulist}
    except Exception:
        return None


def two_site_leg(torch, lib, C, _lib, StiefelFit, T, stream, device, peak_tf, steps):
    """U1 throughput of the 2-site model on a config-4-shaped batch (1024 instances of 9 layers (64, 4)): the factored evaluation
    (csrc/factored.cuh, default) against the block-diagonal D x D chain, device-resident inputs, CUDA events."""
    import numpy as np
    from opentn_b200 import optimization as O
    from opentn_b200 import stiefel as S
    B, K, m, tau = 1024, 16, 9, 0.5
    target = T.exp_operator_dt(T.create_kitaev_liouvillians(N_SITES, D_LOC, 1.0, pbc=True)[0], tau)
    xs0 = np.array([T.super2ortho(x.real, rank=K) for x in O.get_kitaev_trotter_local_ansatz(1.0, tau, (m - 1) // 2)])
    xs = np.empty((B,) + xs0.shape)
    et = np.empty_like(xs)
    for b in range(B):                     # start = Trotter point retracted after a random tangent kick of norm 1e-2
        rng = np.random.default_rng(b)
        kick = np.array([S.project(x, rng.standard_normal(x.shape)) for x in xs0])
        kick *= 1e-2 / np.sqrt(sum(S.canonical_metric(k, k, x) for k, x in zip(kick, xs0)))
        u, _, vt = np.linalg.svd(xs0 + kick, full_matrices=False)
        xs[b] = u @ vt
        et[b] = np.array([S.project(x, rng.standard_normal(x.shape)) for x in xs[b]])
    x = torch.from_numpy(xs).cuda(); e = torch.from_numpy(et).cuda()
    f = torch.empty(B, dtype=torch.float64, device="cuda"); g = torch.empty_like(x); h = torch.empty_like(x)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = {"workload": f"Kitaev gamma=1 N=4 tau={tau} pbc, 2-site ansatz m={m} K={K}, x[B,{m},{4 * K},4], B={B}", "unit": UNIT, "steps": steps}
    for name, kw in (("factored", dict(factored=True)), ("block_diagonal", dict(dense=False))):
        fit = StiefelFit(target, model="local", N=N_SITES, d=D_LOC, metric="canonical", max_batch=B, device=device, **kw)
        hd = fit._handle(m, K, B)
        _lib.check(lib.otn_set_stream(hd.h, C.c_void_p(stream.cuda_stream)))
        step = lambda: _lib.check(lib.otn_triple(hd.h, x.data_ptr(), e.data_ptr(), B, f.data_ptr(), g.data_ptr(), h.data_ptr()))
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        lib.otn_profile_enable(hd.h, 1)
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        pms = (C.c_double * 6)(); pln = (C.c_longlong * 6)(); pfl = (C.c_double * 6)()
        _lib.check(lib.otn_profile_read(hd.h, pms, pln, pfl))
        lib.otn_profile_enable(hd.h, 0)
        ach = pfl[0] / (pms[0] * 1e-3) / 1e12
        out[name] = {"value": B / (ms * 1e-3), "ms_per_step": ms, "dominant_kernel_ms_per_step": pms[0] / steps,
                     "dominant_kernel_tflops_executed": ach, "frac_of_fp64_peak": ach / peak_tf if peak_tf else None,
                     "dense_equivalent_tflops": (9 * (m - 1) * 2 * 256.0 ** 3) * B / (ms * 1e-3) / 1e12}
        fit.close()
    out["speedup_factored_vs_block_diagonal"] = out["factored"]["value"] / out["block_diagonal"]["value"]
    return out


# ---------------------------------------------------------------------------------------------------------------------
# additional legs (all bounded; each returns a dict that becomes a key of the JSON line)
# ---------------------------------------------------------------------------------------------------------------------
TR_KICK = 1e-2


def c3_tr_start():
    """Trotter (Strang) starting point of the C3 workload at Kraus rank 16: [exp(tau/2 L_odd), exp(tau L_even), exp(tau/2 L_odd)] of
    the Kitaev chain (experiments/trust_region_fullsites_ansatz.ipynb cell 1: tau = 4, pbc = False), each layer factorised by
    super2ortho(rank=16) (transformations.py:728-741) on the device (otn_super2ortho).  Set-up, not timed."""
    import numpy as np
    from opentn_b200 import transformations as T
    layers = T.create_trotter_layers(list(T.create_kitaev_liouvillians(N_SITES, D_LOC, 1.0, pbc=False)[:3]), tau=4.0)[1:]
    xs = T.super2ortho_batch(np.stack([np.asarray(x).real for x in layers]), K_RANK)      # device factory (256 x 256 Choi matrices)
    return np.stack([xs[0], xs[1], xs[0]])


def _cpu_tr_c3_worker(args):
    """TR iterations of the oracle (matrix-free Hessian operator) on the SAME instances the device TR leg runs: the kicked Trotter
    starts are handed over as arrays (Kraus-operator signs of a rank-deficient Choi factor are a convention of the factorisation,
    so rebuilding the start on the host would give a different instance)."""
    xs_list, niter = args
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    from oracle import port as P
    T = P.kitaev_fullsites_target(N_SITES, D_LOC, 1.0, 4.0, False)
    spec = P.CostSpec(T, "full")
    t0 = time.perf_counter()
    ncg, fs = 0, []
    for xs in xs_list:
        _, f_it, _, log_ = P.trust_region_fit(spec, list(xs), "canonical", niter=niter)
        ncg += sum(l["n_matvec"] - 1 for l in log_)
        fs.append(f_it)
    return time.perf_counter() - t0, ncg, fs


def _cpu_latency_worker(args):
    """Single-instance trust-region latency of the oracle on configs 1 / 2 (one process, one BLAS thread)."""
    which, niter = args
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    from oracle import port as P
    Li = P.pspl_lindbladians(1.0)
    tau, n = (1.0, 1) if which == "c1" else (0.5, 4)
    T = P.exp_operator_dt(P.create_2local_liouvillians(Li, N_SITES, D_LOC, pbc=True)[0], tau)
    xs0 = [P.super2ortho(x.real, rank=10) for x in P.get_general_trotter_local_ansatz(Li, tau, n=n)]
    spec = P.CostSpec(T, "local", N_SITES, D_LOC)
    t0 = time.perf_counter()
    _, f_it, _, log_ = P.trust_region_fit(spec, xs0, "canonical", niter=niter)
    dt = time.perf_counter() - t0
    return dt / niter, sum(l["n_matvec"] - 1 for l in log_) / niter, list(f_it)


def latency_leg(torch, StiefelFit, device, cpu):
    """Configs 1 and 2 of BASELINE.json: ONE instance (PSPL N=4, K=10; tau=1 n=1 m=3 / tau=0.5 n=4 m=9), Trotter start, canonical
    metric: milliseconds per trust-region iteration of the device solver (otn_tr_run, host arrays in and out) next to the oracle's
    loop on one host core (same start, same iteration count; matrix-free Hessian operator = baseline (ii))."""
    import numpy as np
    from opentn_b200 import optimization as O
    from opentn_b200 import transformations as T
    out = {"unit": "ms per TR iteration (single instance)", "note": "device: StiefelFit.tr_run with host arrays, wall clock incl. "
           "H2D/D2H of x and the report; cpu: oracle/port.py trust_region_fit, 1 process, 1 BLAS thread"}
    Li = [np.kron(T.X, T.I) - np.kron(T.I, T.X), np.kron(T.Z, T.I) - np.kron(T.I, T.Z)]          # PSPL, gamma = 1
    for which, tau, n, niter_dev, niter_cpu in (("c1", 1.0, 1, 5, 5), ("c2", 0.5, 4, 3, 3)):
        target = T.exp_operator_dt(T.create_2local_liouvillians(Li, N_SITES, D_LOC, pbc=True)[0], tau)
        xs0 = np.stack([T.super2ortho(np.asarray(x).real, rank=10) for x in O.get_general_trotter_local_ansatz(Li, tau, n)])
        fit = StiefelFit(target, model="local", N=N_SITES, d=D_LOC, metric="canonical", device=device)
        fit.tr_run(xs0, niter=2)                                   # warm-up: handle, kernel attributes, graph
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, rep = fit.tr_run(xs0, niter=niter_dev)
        torch.cuda.synchronize()
        ms_dev = 1e3 * (time.perf_counter() - t0) / niter_dev
        leg = {"workload": f"PSPL N=4 tau={tau} n={n} (m={2 * n + 1}) K=10, 2-site ansatz, Trotter start, B=1",
               "device_ms_per_iter": ms_dev, "device_niter": niter_dev, "device_n_cg_mean": float(rep["n_cg"].mean()),
               "device_f_first_last": [float(rep["f_iter"][0, 0]), float(rep["f_iter"][-1, 0])]}
        fit.close()
        if cpu is not None:
            s_it, ncg, f_it = cpu.pool.apply(_cpu_latency_worker, ((which, niter_cpu),))
            leg.update({"cpu_ms_per_iter": 1e3 * s_it, "cpu_niter": niter_cpu, "cpu_n_cg_mean": ncg, "cpu_cores": 1,
                        "cpu_f_first_last": [f_it[0], f_it[-1]], "speedup": 1e3 * s_it / ms_dev,
                        "cost_history_max_rel_diff": float(np.max(np.abs(np.array(f_it) - rep["f_iter"][:, 0]) / np.array(f_it)))})
        out[which] = leg
    return out


def _complex_problem(B, K=8, m=3, seed=12):
    """Synthetic complex full-sites instances (N=4: isometries of 16 K x 16, complex128) and a complex target near the model of a
    hidden point; the same on the device and in the oracle worker (seeded)."""
    import numpy as np
    from oracle import complex_port as CP
    rng, rng_x, rng_e = (np.random.default_rng(seed + i) for i in range(3))     # separate streams: instance b does not depend on B
    dim = 16
    T = CP.model([CP.random_stiefel(dim * K, dim, rng) for _ in range(m)])
    T = T + 0.05 * (rng.standard_normal(T.shape) + 1j * rng.standard_normal(T.shape))
    xs = np.array([[CP.random_stiefel(dim * K, dim, rng_x) for _ in range(m)] for _ in range(B)])
    et = np.array([[CP.project(x, rng_e.standard_normal(x.shape) + 1j * rng_e.standard_normal(x.shape)) for x in xb] for xb in xs])
    return T, xs, et


def _cpu_complex_worker(args):
    """(first, count) -> seconds for `count` complex U1 evaluations of the oracle (oracle/complex_port.py), one BLAS thread."""
    first, count = args
    from oracle import complex_port as CP
    T, xs, et = _complex_problem(first + count)
    t0 = time.perf_counter()
    fs = []
    for b in range(first, first + count):
        f, _, _ = CP.triple(T, list(xs[b]), list(et[b]), "canonical")
        fs.append(f)
    return time.perf_counter() - t0, fs


def complex_leg(torch, StiefelFit, device, cpu, B=64, K=8, m=3, steps=5):
    """SURVEY 8f #4, complex isometries: U1 = (f, rgrad, H eta) for B complex N=4 full-sites instances (K = 8, m = 3) through the
    complex handle (chain on the real embedding, upper half of every product computed and mirrored: 4 real D^3 products per complex
    product), device-resident; the complex oracle on the same instances beside it (bounded sample, all cores, one instance each)."""
    import numpy as np
    T, xs, et = _complex_problem(B, K, m)
    xd, ed = torch.from_numpy(xs).cuda(), torch.from_numpy(et).cuda()
    fit = StiefelFit(T, model="full", N=4, d=2, metric="canonical", max_batch=B, device=device)
    outs = fit.triple(xd, ed)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        outs = fit.triple(xd, ed, out=outs)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    f_dev = outs[0].cpu().numpy()
    fit.close()
    D = 256
    out = {"workload": f"complex isometries, N=4 full-sites, m={m}, K={K}, B={B}: U1 per instance; chain on the real embedding (side {2 * D}), 4M products",
           "u1_per_sec": B / (ms * 1e-3), "ms_per_step": ms,
           "chain_tflops_real": 9 * (m - 1) * 4 * 2.0 * D ** 3 * B / (ms * 1e-3) / 1e12}
    if cpu is not None:
        n = min(cpu.cores, B)
        res = cpu.pool.map(_cpu_complex_worker, [(c, 1) for c in range(n)])
        sec = max(r[0] for r in res)
        f_cpu = np.array([r[1][0] for r in res])
        out["cpu"] = {"u1_per_sec": n / sec, "cores": n, "kind": "port", "sample": f"{n} of the {B} instances, one per core, oracle/complex_port.py triple",
                      "cost_max_rel_diff_vs_device": float(np.max(np.abs(f_cpu - f_dev[:n]) / f_cpu))}
        out["speedup_vs_cpu"] = out["u1_per_sec"] / out["cpu"]["u1_per_sec"]
    return out


def host_ceiling_leg(torch, dist, world, x_pin, g_pin, barrier, max_over_ranks, steps):
    """What the host side of the box can move: every rank copies one step's worth of inputs H2D and outputs D2H concurrently
    (two streams, pinned buffers, NO kernels).  The end-to-end U1 rate cannot exceed batch / this time, whatever the GPUs do."""
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    d_in = torch.empty((2,) + tuple(x_pin.shape), dtype=torch.float64, device="cuda")
    d_out = torch.empty_like(d_in)
    h_in = [x_pin, x_pin]; h_out = [g_pin, g_pin]

    def one():
        with torch.cuda.stream(s_in):
            for k in range(2):
                d_in[k].copy_(h_in[k], non_blocking=True)
        with torch.cuda.stream(s_out):
            for k in range(2):
                h_out[k].copy_(d_out[k], non_blocking=True)
    for _ in range(2):
        one()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    torch.cuda.synchronize()
    ms = max_over_ranks(1e3 * (time.perf_counter() - t0))
    barrier()
    nbytes = 4 * x_pin.numel() * 8
    return {"ms_per_step": ms / steps, "gb_per_s_per_rank_both_directions": nbytes * steps / (ms * 1e-3) / 1e9,
            "u1_per_s_ceiling": world * BATCH * steps / (ms * 1e-3),
            "note": "concurrent pinned H2D (x, eta) + D2H (rgrad, heta) of one step per rank, no kernels; wall clock, max over ranks"}


def tr_leg(torch, C, lib, _lib, fit, handle, stream, rank, world, barrier, max_over_ranks, e0, e1, cpu, niter):
    """Unit U2: trust-region iterations/s of the batched device solver on the C3 batch with REALISTIC starts -- the
    Trotter point at Kraus rank 16 retracted after a random tangent kick of canonical norm 1e-2 (seed = global instance index) --
    (a) device-resident x, (b) end to end: pinned host x in, otn_tr_run, host x and report out.  With a CpuPool: the oracle's loop on
    the same instances (first `cores` seeds), same iteration count."""
    import numpy as np
    from opentn_b200 import sweep
    x0 = c3_tr_start()
    seeds = np.arange(rank * BATCH, (rank + 1) * BATCH, dtype=np.uint64)
    xs_dev = sweep.kicked_starts_device(torch.from_numpy(np.broadcast_to(x0, (BATCH,) + x0.shape).copy()).cuda(), seeds, TR_KICK)
    fit.tr_run(xs_dev, niter=1)                                     # warm-up
    barrier()
    e0.record(stream)
    _, rep = fit.tr_run(xs_dev, niter=niter)
    e1.record(stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    out = {"iters_per_sec": world * BATCH * niter / (ms * 1e-3), "unit": "TR iterations/s (all instances)", "niter": niter,
           "starts": f"Trotter point (K=16) + tangent kick of canonical norm {TR_KICK}, seed = global instance index (otn_kicked_starts)",
           "n_cg_mean": float(rep["n_cg"].mean()), "n_cg_max": int(rep["n_cg"].max()), "hvp_total_rank0": int(rep["hvp_total"]),
           "ms_per_iter_batch": ms / niter, "accept_rate": float(rep["accepted"].mean()),
           "f_median_first_last": [float(np.median(rep["f_iter"][0])), float(np.median(rep["f_iter"][-1]))],
           "status_nonzero": int((rep["status"] != 0).sum())}
    # (b) end to end through the C ABI with host arrays: x in (pinned), x out, per-iteration report arrays out
    x_pin = xs_dev.cpu().pin_memory()
    reps = 2
    xws = [x_pin.clone().pin_memory() for _ in range(reps + 1)]      # otn_tr_run works in place: one pinned copy of the starts per call
    f_it = torch.empty((niter, BATCH), dtype=torch.float64).pin_memory()
    ncg = torch.empty((niter, BATCH), dtype=torch.int32).pin_memory()
    status = torch.zeros(BATCH, dtype=torch.int32).pin_memory()
    prm = _lib.TrParams(0.125, 0.01, 0.1, niter, 0, 1e-8, 1e-6)

    def run_e2e(xw):
        rp = _lib.TrReport(f_it.data_ptr(), None, None, ncg.data_ptr(), None, None, None, None, status.data_ptr(), 0)
        _lib.check(lib.otn_tr_run(handle.h, xw.data_ptr(), BATCH, C.byref(prm), C.byref(rp)))
    run_e2e(xws[reps])
    barrier()
    t0 = time.perf_counter()
    for r_ in range(reps):
        run_e2e(xws[r_])
    torch.cuda.synchronize()
    ms_e2e = max_over_ranks(1e3 * (time.perf_counter() - t0)) / reps
    barrier()
    per = x_pin.numel() * 8
    out["e2e"] = {"iters_per_sec": world * BATCH * niter / (ms_e2e * 1e-3), "ms_per_call": ms_e2e, "niter_per_call": niter,
                  "h2d_bytes_per_call": per, "d2h_bytes_per_call": per + niter * BATCH * 12 + BATCH * 4,
                  "mode": "otn_tr_run with pinned host arrays: x in, x + f_iter + n_cg + status out (in place); wall clock around the calls",
                  "matches_device_resident": bool(np.allclose(f_it.numpy(), rep["f_iter"], rtol=1e-12))}
    if cpu is not None:
        n_cpu = cpu.cores
        xs_host = x_pin.numpy()
        res = cpu.pool.map(_cpu_tr_c3_worker, [([xs_host[c]], niter) for c in range(n_cpu)])
        wall = max(r[0] for r in res)
        f_cpu = np.array([r[2][0] for r in res]).T                     # [niter][n_cpu]
        out["cpu"] = {"iters_per_sec": n_cpu * niter / wall, "cores": n_cpu, "kind": "port",
                      "sample": f"{n_cpu} of the same instances (1 per process, 1 BLAS thread), {niter} TR iterations each, {wall:.1f} s",
                      "n_cg_mean": sum(r[1] for r in res) / (n_cpu * niter),
                      "cost_history_max_rel_diff_vs_device": float(np.max(np.abs(f_cpu - rep["f_iter"][:, :n_cpu]) / f_cpu))}
        out["speedup_vs_cpu"] = out["iters_per_sec"] / out["cpu"]["iters_per_sec"]
    return out


C4_TAUS, C4_RANKS, C4_STEPS = [0.25, 0.5, 1.0, 2.0], [2, 4, 8, 16], 4


def _cpu_c4_worker(args):
    """One fit of config 4 on the oracle, from the kicked start the device run used (handed over as an array): seconds for
    `niter` TR iterations and the first cost."""
    xs, ti, niter = args
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    from oracle import port as P
    Lk = P.get_kitaev_nn_linbladian(1.0)
    T = P.exp_operator_dt(P.create_2local_liouvillians(Lk, N_SITES, D_LOC, pbc=True)[0], C4_TAUS[ti])
    spec = P.CostSpec(T, "local", N_SITES, D_LOC)
    t0 = time.perf_counter()
    _, f_it, _, _ = P.trust_region_fit(spec, list(xs), "canonical", niter=niter)
    return time.perf_counter() - t0, f_it[0]


def c4_leg(torch, dist, rank, world, device, barrier, max_over_ranks, cpu, seeds_per_cell=256, niter=5):
    """BASELINE config 4: tau x rank x seed sweep, 4 taus x ranks {2,4,8,16} x `seeds_per_cell` seeds = 4096 instances of the 2-site
    model (N=4, m=9), `niter` TR iterations each, STRONG scaling: the 16 (K, tau) cells are split over the ranks
    (sweep.plan_cells), one final all-gather.  Reported: fits/s over the whole job, set-up (device Liouvillian, expm sweep, Choi
    factorisation, kicked starts) and solve separately."""
    import numpy as np
    from opentn_b200 import sweep
    from opentn_b200 import transformations as T
    Lnn = [T.get_kitaev_nn_linbladian(1.0)]
    t_fit = time.perf_counter()
    fit = sweep.make_fit(Lnn, N_SITES, D_LOC, C4_TAUS, device=device)               # targets: device Liouvillian + batched expm
    torch.cuda.synchronize()
    t_fit = time.perf_counter() - t_fit
    # warm-up with the full batch sizes: creates the handles (work buffers) of every isometry shape and configures the kernels
    sweep.run_sweep(Lnn, N_SITES, D_LOC, C4_TAUS, C4_RANKS, list(range(seeds_per_cell)), n_steps=C4_STEPS, niter=1, device=device,
                    rank=rank, world=world, gather=world > 1, fit=fit, to_host=False)
    barrier()
    tm = {}
    t0 = time.perf_counter()
    res = sweep.run_sweep(Lnn, N_SITES, D_LOC, C4_TAUS, C4_RANKS, list(range(seeds_per_cell)), n_steps=C4_STEPS, niter=niter,
                          device=device, rank=rank, world=world, gather=world > 1, timings=tm, to_host=False, fit=fit)
    torch.cuda.synchronize()
    wall = max_over_ranks(time.perf_counter() - t0)
    fit.close()
    setup_s, solve_s, gather_s = (max_over_ranks(tm.get(k, 0.0)) for k in ("setup_s", "solve_s", "gather_s"))
    barrier()
    n_inst = len(C4_TAUS) * len(C4_RANKS) * seeds_per_cell
    x0_dev = {K: r.get("x0") for K, r in res.items()}                 # starts stay on the device (only the CPU leg reads a few)
    res = {K: {k: v.cpu().numpy() for k, v in r.items() if k not in ("x0", "x_final")} for K, r in res.items()}
    got = sum(len(r["f0"]) for r in res.values())
    improved = float(np.mean(np.concatenate([r["f_final"] < r["f0"] for r in res.values()])))
    out = {"workload": f"Kitaev gamma=1 N=4 pbc, 2-site ansatz n=4 (m=9): taus {C4_TAUS} x K {C4_RANKS} x {seeds_per_cell} seeds = {n_inst} fits, "
                       f"{niter} TR iterations each, Trotter start + 1e-2 kick",
           "scaling": "strong", "fits_per_sec": n_inst / wall, "tr_iters_per_sec": n_inst * niter / wall, "wall_s": wall,
           "setup_s": setup_s, "solve_s": solve_s, "gather_s": gather_s,
           "instances_returned_rank0": got, "improved_fraction": improved,
           "status_nonzero": int(sum((r["status"] != 0).sum() for r in res.values())),
           "target_setup_s": t_fit,
           "partition": "K-major (K, tau) cell list dealt to the ranks in a snake (cheap and expensive cells paired); ONE packed "
                        "all-gather of the whole sweep at the end; handles created by a warm-up sweep (not timed), targets by make_fit "
                        "(target_setup_s, not in wall_s)"}
    if cpu is not None and world == 1:
        picks = [(C4_RANKS[c % 4], (c // 4) % 4, c % seeds_per_cell) for c in range(cpu.cores)]     # (K, tau index, seed) cycling
        jobs, dev_f0 = [], []
        for K, ti, sd in picks:
            r = res[K]
            i = int(np.where((r["tau_index"] == ti) & (r["seed"] == sd))[0][0])
            jobs.append((x0_dev[K][i].cpu().numpy(), ti, 1))
            dev_f0.append(r["f0"][i])
        resc = cpu.pool.map(_cpu_c4_worker, jobs)
        wall_c = max(r[0] for r in resc)
        chk = [abs(d - c[1]) / c[1] for d, c in zip(dev_f0, resc)]       # same instance: first costs agree to rounding
        out["cpu"] = {"tr_iters_per_sec": cpu.cores / wall_c, "cores": cpu.cores, "kind": "port",
                      "sample": f"{cpu.cores} instances of the sweep (K and tau cycling), 1 TR iteration each, {wall_c:.1f} s",
                      "first_cost_max_rel_diff_vs_device": float(max(chk)) if chk else None}
        out["speedup_vs_cpu"] = out["tr_iters_per_sec"] / out["cpu"]["tr_iters_per_sec"]
    return out



def _cpu_zgemm_worker(D):
    """The reference's composition step on the host: one complex128 D x D product with every core (numpy / BLAS threads), the
    operation timed in experiments/gradient.ipynb cells 195-197 (9.4-15 s on the author's machine)."""
    import numpy as np
    rng = np.random.default_rng(0)
    A = rng.standard_normal((D, D)) + 1j * rng.standard_normal((D, D))
    B = rng.standard_normal((D, D)) + 1j * rng.standard_normal((D, D))
    t0 = time.perf_counter()
    A @ B
    return time.perf_counter() - t0


def c5_leg(torch, dist, rank, world, device, barrier, max_over_ranks, with_cpu):
    """BASELINE config 5: one synthetic N=6 instance (D = 4096), three complex layer superoperators (rank-16 two-site layers embedded
    to 4096 x 4096, complex128 as (re, im) planes), the composition Y3 Y2 Y1 = 2 complex = 8 real D^3 products (4M) on FP64 DMMA, rows of
    every product sharded over the ranks; exchange fused into the GEMM epilogue (NVLink peer stores + device-side flags) and, for
    comparison, by NCCL all-gather.  STRONG scaling: the efficiency is measured inside the run against rank 0 computing the same
    products alone."""
    import numpy as np
    from opentn_b200 import transformations as T
    from opentn_b200.rowsharded import RowShardedChain
    T.set_device(device)
    D, m, K = 4096, 3, 16
    rng = np.random.default_rng(5)
    planes = []
    for l in range(m):                      # same layers on every rank
        pair = []
        for part in range(2):
            q, _ = np.linalg.qr(rng.standard_normal((4 * K, 4)))
            pair.append(torch.from_numpy(T.local2full(T.ortho2super(q), 6, 2, shift_pbc=bool(l % 2))).cuda() * (1.0 if part == 0 else 0.1))
        planes.append(tuple(pair))
    flops = 4 * (m - 1) * 2.0 * D ** 3          # (m - 1) complex products, 4 real D^3 products each (4M)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(chain, reps=3):
        chain.compose_complex(planes)
        barrier()
        e0.record()
        for _ in range(reps):
            out = chain.compose_complex(planes)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / reps, out

    # single-GPU time of the same products (every rank computes everything: world = 1 view of the chain)
    solo = RowShardedChain(D, device=device, single=True)
    ms_solo, M_solo = timed(solo)
    out = {"workload": f"synthetic N=6 (D={D}) complex layers, m={m}, composition Y3 Y2 Y1 = {m - 1} complex = {4 * (m - 1)} real D^3 products (4M), "
                       f"rows sharded over {world} GPU(s)",
           "scaling": "strong", "unit": "TFLOP/s (real flops, aggregate)", "flops_per_composition": flops,
           "single_gpu": {"ms": ms_solo, "tflops": flops / ms_solo / 1e9}}
    if world > 1:
        for name, p2p in (("fused_peer_store", True), ("nccl_allgather", False)):
            ch = RowShardedChain(D, p2p=p2p, nbuf=8)
            ms, M = timed(ch)
            same = bool(torch.equal(M[0], M_solo[0]) and torch.equal(M[1], M_solo[1]))
            out[name] = {"ms": ms, "tflops": flops / ms / 1e9, "efficiency_vs_single_gpu": ms_solo / (world * ms),
                         "bitwise_equal_to_single_gpu": same}
            ch.close()
        out["value_tflops"] = out["fused_peer_store"]["tflops"]
    else:
        out["value_tflops"] = out["single_gpu"]["tflops"]
    # parity spot check against NumPy on 4 columns of the result (the full product on the host takes ~10 s per product)
    if rank == 0:
        cols = [0, 1023, 2048, 4095]
        Yc = [(pr[0].cpu().numpy() + 1j * pr[1].cpu().numpy()) for pr in planes]
        v = Yc[0][:, cols]
        for Y in Yc[1:]:
            v = Y @ v
        got = M_solo[0][:, cols].cpu().numpy() + 1j * M_solo[1][:, cols].cpu().numpy()
        out["max_rel_err_vs_numpy_4_columns"] = float(np.abs(got - v).max() / np.abs(v).max())
        if with_cpu:
            t = _cpu_zgemm_worker(D)
            out["cpu"] = {"seconds_per_complex_product": t, "tflops_real_equivalent": 8.0 * D ** 3 / t / 1e12, "cores": os.cpu_count(),
                          "kind": "port", "sample": "one complex128 4096^3 product, numpy @ with all BLAS threads (transformations.py:850 on the host)"}
            out["speedup_vs_cpu"] = out["value_tflops"] / out["cpu"]["tflops_real_equivalent"]
    torch.cuda.empty_cache()
    out["fit"] = c5_fit_leg(torch, dist, rank, world, device, barrier, max_over_ranks)
    return out


def c5_fit_leg(torch, dist, rank, world, device, barrier, max_over_ranks):
    """Config 5 as a FIT: one dense N=6 instance (2-site ansatz, K = 16, m = 3 layers embedded to 4096 x 4096, Kitaev pbc target
    exp(4L)), the whole unit of work U1 = (f, rgrad, H eta) and one trust-region iteration through a row-sharded handle
    (StiefelFit.row_shard -> otn_rowshard_attach): every one of the 18 chain products of a U1 is split by rows, tiles stored into
    all peers inside the GEMM, device-side barriers.  Efficiency against the un-sharded dense handle on one GPU, same run; results
    bitwise equal.  (The factored evaluation does the same U1 in ~1.5 ms on one GPU without forming a D x D matrix: this leg is
    about dense layers, the reference's formulation.)"""
    import numpy as np
    from opentn_b200 import StiefelFit
    from opentn_b200 import transformations as T
    T.set_device(device)
    N, d, m, K, D = 6, 2, 3, 16, 4096
    full = T.create_kitaev_liouvillians(N, d, 1.0, pbc=True, backend="cuda")[0]
    target = np.ascontiguousarray(T.exp_operator_dt(full, 4.0, backend="cuda").real)
    rng = np.random.default_rng(6)                              # same instance on every rank
    xs = np.stack([np.linalg.qr(rng.standard_normal((4 * K, 4)))[0] for _ in range(m)])[None]
    et = rng.standard_normal(xs.shape)
    et = et - xs @ (0.5 * (np.swapaxes(xs, -1, -2) @ et + np.swapaxes(et, -1, -2) @ xs))     # tangent projection (stiefel.py:23-30)
    xd, ed = torch.from_numpy(xs).cuda(), torch.from_numpy(et).cuda()
    flops_u1 = 18 * 2.0 * D ** 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fit, reps=3):
        outs = fit.triple(xd, ed)
        barrier()
        e0.record()
        for _ in range(reps):
            outs = fit.triple(xd, ed, out=outs)
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1)) / reps
        barrier()
        t0 = time.perf_counter()
        _, rep = fit.tr_run(xd, niter=1)
        torch.cuda.synchronize()
        barrier()
        tr_ms = max_over_ranks(1e3 * (time.perf_counter() - t0))
        return ms, outs, tr_ms, rep

    solo = StiefelFit(target, model="local", N=N, d=d, metric="canonical", dense=True, device=device)
    solo.triple(xd, ed)
    stream = torch.cuda.current_stream()
    ms1, o1, tr1, rep1 = timed(solo)
    # where the single-GPU time goes (per-launch CUDA events; the kernels outside the chain are replicated on every rank)
    import ctypes as C
    from opentn_b200 import _lib
    lib = _lib.load()
    (hd,) = solo._handles.values()
    lib.otn_profile_enable(hd.h, 1)
    solo.triple(xd, ed, out=o1)
    pms = (C.c_double * 6)(); pln = (C.c_longlong * 6)(); pfl = (C.c_double * 6)()
    _lib.check(lib.otn_profile_read(hd.h, pms, pln, pfl))
    lib.otn_profile_enable(hd.h, 0)
    breakdown = {"chain_products_ms": pms[0], "kraus_assembly_and_embedding_ms": pms[1], "back_embedding_and_pull_back_ms": pms[2],
                 "riemannian_epilogue_ms": pms[3], "launches": int(sum(pln))}
    solo.close()
    out = {"workload": f"dense N=6 2-site fit (D={D}, m={m}, K={K}): U1 = cost + Riemannian gradient + HVP = 18 real D^3 products, and one TR iteration",
           "flops_per_u1": flops_u1,
           "single_gpu": {"u1_ms": ms1, "tflops": flops_u1 / ms1 / 1e9, "tr_iteration_ms": tr1, "n_cg": int(rep1["n_cg"][0, 0]), "cost": float(o1[0][0]),
                          "profiled_u1_ms_by_category": breakdown}}
    if world > 1:
        fit = StiefelFit(target, model="local", N=N, d=d, metric="canonical", dense=True, device=device).row_shard()
        ms, o, trm, rep = timed(fit)
        same = bool(torch.equal(o[0], o1[0]) and torch.equal(o[1], o1[1]) and torch.equal(o[2], o1[2]))
        rel = max(float((o[i] - o1[i]).abs().max() / o1[i].abs().max()) for i in range(3))
        # where the sharded time goes: per-launch events of this rank's kernels (the chain products include their peer stores;
        # the device-side barriers are not in any category), once more with the peer stores switched off (numbers then garbage)
        (hs,) = fit._handles.values()
        diag = {}
        for name, env in (("with_peer_stores", None), ("without_peer_stores", "1")):
            if env:
                os.environ["OTN_RS_NO_PEER_STORES"] = env
            try:
                scratch = tuple(t.clone() for t in o)
                fit.triple(xd, ed, out=scratch)
                barrier()
                lib.otn_profile_enable(hs.h, 1)
                t0 = time.perf_counter()
                fit.triple(xd, ed, out=scratch)
                torch.cuda.synchronize()
                wall = 1e3 * (time.perf_counter() - t0)
                _lib.check(lib.otn_profile_read(hs.h, pms, pln, pfl))
                lib.otn_profile_enable(hs.h, 0)
                barrier()
            finally:
                os.environ.pop("OTN_RS_NO_PEER_STORES", None)
            diag[name] = {"u1_wall_ms_profiled": max_over_ranks(wall), "chain_products_ms": max_over_ranks(pms[0]),
                          "replicated_kernels_ms": max_over_ranks(pms[1] + pms[2] + pms[3])}
        out["row_sharded"] = {"u1_ms": ms, "tflops": flops_u1 / ms / 1e9, "efficiency_vs_single_gpu": ms1 / (world * ms),
                              "tr_iteration_ms": trm, "tr_efficiency_vs_single_gpu": tr1 / (world * trm), "n_cg": int(rep["n_cg"][0, 0]),
                              "bitwise_equal_to_single_gpu": same, "max_rel_diff_vs_single_gpu": rel,
                              "tr_cost_rel_diff": float(np.max(np.abs(rep["f_iter"] - rep1["f_iter"]) / np.abs(rep1["f_iter"]))),
                              "barriers_ok": bool(fit.row_shard_ok()), "diagnostics": diag}
        fit.close()
    del stream
    return out


def run_gpu(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    from opentn_b200 import StiefelFit, _lib
    from opentn_b200 import transformations as T

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(local_rank)      # before any pinned allocation: host staging buffers land next to the GPU
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _lib.load()

    if args.workload in ("c4", "c5"):           # only that configuration's leg, as the line
        def barrier_():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        def max_over_ranks_(v):
            if world == 1:
                return v
            t = torch.tensor([v], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        cpu_pool = CpuPool() if (not args.no_cpu and world == 1) else None
        try:
            if args.workload == "c4":
                leg = c4_leg(torch, dist, rank, world, local_rank, barrier_, max_over_ranks_, cpu_pool)
                line = {"metric": "trust-region fits/sec (config 4: tau x rank x seed sweep, 4096 fits x 5 iterations)", "value": leg["fits_per_sec"],
                        "unit": "fits/s", "scaling": "strong", "ms_per_step": 1e3 * leg["wall_s"]}
            else:
                leg = c5_leg(torch, dist, rank, world, local_rank, barrier_, max_over_ranks_, with_cpu=cpu_pool is not None)
                line = {"metric": "row-sharded N=6 complex composition (config 5), real TFLOP/s aggregate", "value": leg["value_tflops"],
                        "unit": "TFLOP/s", "scaling": "strong",
                        "ms_per_step": leg.get("fused_peer_store", leg["single_gpu"])["ms"]}
        finally:
            if cpu_pool is not None:
                cpu_pool.close()
        line.update({"n_gpus": world, "steps": 1, "warmup": 1, "higher_is_better": True, "vs_baseline": None, "dtype": "f64",
                     "data": "synthetic", "config": {"workload": leg["workload"]}, args.workload: leg})
        if rank == 0:
            print(json.dumps(line))
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- set-up (not timed): target, inputs -----------------------------------------------------------------------
    Lvec = T.create_kitaev_liouvillians(N_SITES, D_LOC, 1.0, pbc=False)[0]
    target = T.exp_operator_dt(Lvec, 4.0)
    xs_h, et_h = make_inputs(rank * BATCH, BATCH)          # weak scaling: every rank owns its own B instances
    x_pin = torch.from_numpy(xs_h).pin_memory()
    e_pin = torch.from_numpy(et_h).pin_memory()
    x_dev = x_pin.cuda(non_blocking=True)
    e_dev = e_pin.cuda(non_blocking=True)
    f_dev = torch.empty(BATCH, dtype=torch.float64, device="cuda")
    g_dev = torch.empty_like(x_dev)
    h_dev = torch.empty_like(x_dev)
    f_pin = torch.empty(BATCH, dtype=torch.float64).pin_memory()
    g_pin = torch.empty_like(x_pin).pin_memory()
    h_pin = torch.empty_like(x_pin).pin_memory()
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def make_fit(dense):
        fit_ = StiefelFit(target, model="full", N=N_SITES, d=D_LOC, metric="canonical", max_batch=BATCH, device=local_rank, dense=dense)
        handle_ = fit_._handle(M_LAYERS, K_RANK, BATCH)
        _lib.check(lib.otn_set_stream(handle_.h, C.c_void_p(stream.cuda_stream)))   # torch.cuda.Event then brackets our kernels
        return fit_, handle_

    def timed_device_steps(handle_, steps, min_warm_s):
        """>= W warm-up steps (and >= min_warm_s under load), then two timed regions of `steps` steps of otn_triple on
        device-resident inputs: (1) the plain one (`value`); (2) the same steps with every kernel launch bracketed by a CUDA-event
        pair on its stream (otn_profile_*; the roofline numbers).  Per-launch timing needs launches that do not overlap, so in (2)
        the library keeps the cotangent pull-backs on the main stream instead of its back stream: (2) is slightly slower."""
        def step():
            _lib.check(lib.otn_triple(handle_.h, x_dev.data_ptr(), e_dev.data_ptr(), BATCH, f_dev.data_ptr(), g_dev.data_ptr(), h_dev.data_ptr()))
        t_warm = time.perf_counter()
        n_warm_ = 0
        while n_warm_ < max(args.warmup, 3) or time.perf_counter() - t_warm < min_warm_s:
            step()
            n_warm_ += 1
        barrier()
        launches0 = lib.otn_launch_count()
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        barrier()
        ms_total_ = max_over_ranks(e0.elapsed_time(e1))
        launches_ = lib.otn_launch_count() - launches0
        lib.otn_profile_enable(handle_.h, 1)
        step()
        barrier()
        pms_ = (C.c_double * 6)(); pln_ = (C.c_longlong * 6)(); pfl_ = (C.c_double * 6)()
        _lib.check(lib.otn_profile_read(handle_.h, pms_, pln_, pfl_))     # discard the records of the untimed step
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        barrier()
        ms_prof_ = max_over_ranks(e0.elapsed_time(e1))
        _lib.check(lib.otn_profile_read(handle_.h, pms_, pln_, pfl_))
        lib.otn_profile_enable(handle_.h, 0)
        return ms_total_, launches_, list(pms_), list(pln_), list(pfl_), n_warm_, ms_prof_

    peak_tf = measure_fp64_peak(torch) if rank == 0 else None

    # ---- device-resident timing (`value`): default (block-diagonal) evaluation ------------------------------------------
    fit, handle = make_fit(dense=False)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()          # nvidia-smi samples every 100 ms; "under load" samples are selected by power draw
    ms_total, launches, pms, pln, pfl, n_warm, ms_prof = timed_device_steps(handle, args.steps, args.min_warm_s)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms_total / args.steps
    value = world * BATCH * args.steps / (ms_total * 1e-3)

    # ---- end-to-end timing: pinned host buffers through the C ABI, H2D + D2H inside the timed region -----------------
    # (a) synchronous: one otn_triple call per step, each returning with its results on the host (per-call latency);
    # (b) streamed (`e2e.value`): otn_triple_submit / otn_triple_wait, step i+1 submitted before step i is waited for, two
    #     rotating sets of pinned host arrays (set B holds the instances in reverse order, so consecutive steps move
    #     different data).  Every step still copies its 201 MB of inputs in and its 201 MB of results out inside the timed
    #     region; only the exposed head / tail copies of a step now hide behind the kernels of its neighbours.
    def step_e2e():
        _lib.check(lib.otn_triple(handle.h, x_pin.data_ptr(), e_pin.data_ptr(), BATCH, f_pin.data_ptr(), g_pin.data_ptr(), h_pin.data_ptr()))

    for _ in range(3):
        step_e2e()
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step_e2e()
    e1.record(stream)
    barrier()
    ms_e2e_sync = max_over_ranks(e0.elapsed_time(e1))
    f_sync = f_pin.clone()

    x_pin2 = torch.flip(x_pin, dims=[0]).contiguous().pin_memory()
    e_pin2 = torch.flip(e_pin, dims=[0]).contiguous().pin_memory()
    f_pin2 = torch.empty_like(f_pin).pin_memory()
    g_pin2 = torch.empty_like(g_pin).pin_memory()
    h_pin2 = torch.empty_like(h_pin).pin_memory()
    sets = [(x_pin, e_pin, f_pin, g_pin, h_pin), (x_pin2, e_pin2, f_pin2, g_pin2, h_pin2)]

    def stream_steps(n):
        tickets = []
        for i in range(n):
            xs_, es_, fs_, gs_, hs_ = sets[i & 1]
            t = C.c_longlong(-1)
            _lib.check(lib.otn_triple_submit(handle.h, xs_.data_ptr(), es_.data_ptr(), BATCH, fs_.data_ptr(), gs_.data_ptr(),
                                             hs_.data_ptr(), C.byref(t)))
            tickets.append(t.value)
            if i >= 1:
                _lib.check(lib.otn_triple_wait(handle.h, tickets[i - 1]))   # results of step i-1 are on the host: its set is free
        _lib.check(lib.otn_triple_wait(handle.h, tickets[-1]))

    for t_ in (f_pin, g_pin, h_pin):
        t_.zero_()
    stream_steps(max(args.warmup, 3))
    barrier()
    e0.record(stream)
    stream_steps(args.steps)
    e1.record(stream)
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = world * BATCH * args.steps / (ms_e2e * 1e-3)
    e2e_sync_value = world * BATCH * args.steps / (ms_e2e_sync * 1e-3)
    h2d = 2 * x_pin.numel() * 8
    d2h = 2 * x_pin.numel() * 8 + BATCH * 8
    assert torch.equal(f_sync, f_pin) and torch.equal(torch.flip(f_pin2, dims=[0]), f_pin), "streamed and synchronous host paths disagree"
    assert torch.equal(torch.flip(g_pin2, dims=[0]), g_pin) and torch.equal(torch.flip(h_pin2, dims=[0]), h_pin)

    # ---- device-resident, streamed: the same submissions with device arrays (no copies).  What changes against the blocking
    # otn_triple loop above is only the step boundary: no host synchronisation between steps, chunks alternating between the two
    # compute streams, so the head of step i+1 runs under the tail of step i.  Two rotating output sets.
    f_dev2, g_dev2, h_dev2 = torch.empty_like(f_dev), torch.empty_like(g_dev), torch.empty_like(h_dev)
    dev_sets = [(f_dev, g_dev, h_dev), (f_dev2, g_dev2, h_dev2)]

    def dev_stream_steps(n):
        tickets = []
        for i in range(n):
            fs_, gs_, hs_ = dev_sets[i & 1]
            t = C.c_longlong(-1)
            _lib.check(lib.otn_triple_submit(handle.h, x_dev.data_ptr(), e_dev.data_ptr(), BATCH, fs_.data_ptr(), gs_.data_ptr(),
                                             hs_.data_ptr(), C.byref(t)))
            tickets.append(t.value)
            if i >= 1:
                _lib.check(lib.otn_triple_wait(handle.h, tickets[i - 1]))
        _lib.check(lib.otn_triple_wait(handle.h, tickets[-1]))

    g_blocking = g_dev.clone()
    h_blocking = h_dev.clone()
    dev_stream_steps(max(args.warmup, 3))
    barrier()
    launches0 = lib.otn_launch_count()
    e0.record(stream)
    dev_stream_steps(args.steps)
    e1.record(stream)
    barrier()
    ms_dev_stream = max_over_ranks(e0.elapsed_time(e1))
    launches_stream = lib.otn_launch_count() - launches0
    assert torch.equal(g_dev, g_blocking) and torch.equal(g_dev2, g_blocking) and torch.equal(h_dev2, h_blocking), \
        "streamed and blocking device-resident paths disagree"
    value_streamed = world * BATCH * args.steps / (ms_dev_stream * 1e-3)

    # ---- parity spot check on the timed data (device vs host path) and final gather (the only collective) -------------
    assert torch.allclose(f_dev.cpu(), f_pin, rtol=1e-13, atol=0), "device-resident and host-staged paths disagree"
    gather_ms = None
    if world > 1:
        allf = [torch.empty_like(f_dev) for _ in range(world)]
        barrier()
        e0.record(stream)
        dist.all_gather(allf, f_dev)
        e1.record(stream)
        barrier()
        gather_ms = max_over_ranks(e0.elapsed_time(e1))

    # ---- host-side ceiling of the end-to-end path: concurrent H2D + D2H of one step per rank, no kernels ---------------------
    host_ceiling = host_ceiling_leg(torch, dist, world, x_pin, g_pin, barrier, max_over_ranks, max(5, args.steps // 5))

    # ---- CPU pool for the legs that time the oracle beside the device (N = 1 only: rank 0 would compete with N-1 ranks) ---------
    cpu_pool = None
    if not args.no_cpu and world == 1:
        os.sched_setaffinity(0, all_cpus)  # the CPU baseline gets every core of the box, not just the GPU's NUMA node
        cpu_pool = CpuPool()

    # ---- TR iterations / s (unit U2): kicked Trotter starts, device-resident and end to end, oracle on the same instances -------
    tr = None
    if not args.no_tr:
        tr = tr_leg(torch, C, lib, _lib, fit, handle, stream, rank, world, barrier, max_over_ranks, e0, e1, cpu_pool, niter=4)
    fit.close()

    # ---- configs 1 / 2: single-instance latency; config 4: the sharded sweep ------------------------------------------------------
    latency = None
    if not args.no_latency and rank == 0 and world == 1:
        latency = latency_leg(torch, StiefelFit, local_rank, cpu_pool)
    cplx = None
    if not args.no_complex and rank == 0 and world == 1:
        torch.cuda.empty_cache()
        cplx = complex_leg(torch, StiefelFit, local_rank, cpu_pool)
    c4 = None
    if not args.no_c4:
        torch.cuda.empty_cache()
        c4 = c4_leg(torch, dist, rank, world, local_rank, barrier, max_over_ranks, cpu_pool)
    c5 = None
    if not args.no_c5:
        torch.cuda.empty_cache()
        c5 = c5_leg(torch, dist, rank, world, local_rank, barrier, max_over_ranks, with_cpu=(not args.no_cpu and world == 1))

    # ---- same workload through the dense D x D formulation (what the reference executes; SURVEY 8d flop count) -----------
    dense_raw = None
    if not args.no_dense:
        torch.cuda.empty_cache()
        fit_d, handle_d = make_fit(dense=True)
        d_steps = max(3, args.steps // 2)
        dense_raw = timed_device_steps(handle_d, d_steps, min(0.2, args.min_warm_s)) + (d_steps,)
        fit_d.close()

    # ---- 2-site model (BASELINE config 4 shape: Kitaev pbc, tau = 0.5, m = 9, K = 16): factored vs block-diagonal chain ------
    two_site = None
    if not args.no_local and rank == 0 and world == 1:
        torch.cuda.empty_cache()
        two_site = two_site_leg(torch, lib, C, _lib, StiefelFit, T, stream, local_rank, peak_tf, max(5, args.steps // 5))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (chain GEMM on the FP64 tensor pipe) -------------------------------------------
    def roofline_of(pms_, pln_, pfl_, ms_total_, steps_):
        gemm_ms, gemm_n, gemm_fl = pms_[0], pln_[0], pfl_[0]
        ach = gemm_fl / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
        return {"achieved": ach, "frac": (ach / peak_tf) if ach and peak_tf else None, "launches_timed": int(gemm_n),
                "avg_launch_ms": gemm_ms / gemm_n if gemm_n else None,
                "executed_flops_per_launch": gemm_fl / gemm_n if gemm_n else None,
                "profiled_ms_per_step": ms_total_ / steps_, "kernel_share_of_step": gemm_ms / ms_total_,
                "other_kernels_ms_per_step": {"assembly": pms_[1] / steps_, "back_assembly": pms_[2] / steps_,
                                              "riemannian_epilogue": pms_[3] / steps_, "other": pms_[5] / steps_}}

    peak_file = None
    try:
        peak_file = json.load(open(os.path.join(ROOT, "profiles", "fp64_peaks_r01.json")))["dgemm_8192"] / 1e3
    except Exception:
        pass
    peak_source = ("cuBLAS DGEMM 8192^3 via torch.matmul, best of 3, measured in this run (MEASURED_PEAKS.json has no FP64 figure; "
                   "DMMA issue-rate ceiling 148 SM x 128 flop/clk x 1.965 GHz = 37.2 TFLOP/s)")
    roofline = {"bound": "tensor", "kernel": "chain_fused_kernel<144,72,24,4> + chain_fused_kernel<128,32,32,4> (FP64 DMMA.8x8x4): one CTA per "
                                             "instance and diagonal block runs all 18 chain products of a U1 (csrc/chain_fused.cuh); one launch per block and "
                                             "pipeline chunk (512 instances), the two blocks' launches run concurrently and are timed as one group",
                "peak": peak_tf, "unit": "TFLOP/s", "traffic": 5.51e9,
                "traffic_note": "dram read+write of one timed group (both blocks, 512 instances), from ncu --set full of the 1024-instance "
                                "launches (profiles/ncu_step_final_r02.txt: 144-block 4.66 GB read + 1.80 GB written, 128-block 3.16 + 1.39 GB), halved.  "
                                "Operand bytes of the 18 products counted once per product (36 reads + 12 writes of 136^2*8 + 120^2*8 B per "
                                "instance) = 6.5e9 B per group: the re-reads of intermediates partly hit L2; the kernels are tensor-pipe bound "
                                "(85.5 % / 88.7 % active, DRAM 26-28 %)",
                "peak_source": peak_source, "peak_committed_r01": peak_file}
    roofline.update(roofline_of(pms, pln, pfl, ms_prof, args.steps))
    roofline["note"] = ("launches_timed counts fork..join groups (both blocks' launches of a step run concurrently on two streams and are timed as one group). "
                        "`achieved` counts EXECUTED flops of the block-diagonal evaluation, which since the fused kernel skips the padding "
                        "tiles of the 144 / 128 blocks are exactly the USEFUL flops 2*(136^3 + 120^3) per product (round 1 executed and counted the "
                        "padded 2*(144*144*136 + 128*128*120), 12.7 % more). Dense-equivalent rate (SURVEY 8d algorithmic count, 2*256^3 per "
                        "product) is `step_tflops_dense_equivalent`; the dense formulation itself is measured in `dense_formulation`.")
    roofline["achieved_if_padding_counted"] = (roofline["achieved"] * (144 * 144 * 136 + 128 * 128 * 120) / (136.0 ** 3 + 120.0 ** 3)
                                               if roofline.get("achieved") else None)
    roofline["step_tflops_dense_equivalent"] = FLOPS_PER_TRIPLE * BATCH / (ms_per_step * 1e-3) / 1e12

    dense_leg = None
    if dense_raw is not None:
        ms_d, launches_d, pms_d, pln_d, pfl_d, _, ms_prof_d, d_steps = dense_raw
        rf = {"bound": "tensor", "kernel": "gemm_dmma_kernel<64,64,16,32,32,3> (FP64 DMMA.8x8x4), 2*256^3 per product",
              "peak": peak_tf, "unit": "TFLOP/s", "traffic": 1.59e9,
              "traffic_note": "dram read+write per launch of a single-product GEMM over 1024 instances from ncu --set full "
                              "(profiles/ncu_gemm_dense_r01.txt); algorithmic 3*D^2*8*1024 = 1.61e9 B"}
        rf.update(roofline_of(pms_d, pln_d, pfl_d, ms_prof_d, d_steps))
        dense_leg = {"value": world * BATCH * d_steps / (ms_d * 1e-3), "unit": UNIT, "ms_per_step": ms_d / d_steps, "steps": d_steps,
                     "gpu_launches": int(launches_d), "roofline": rf,
                     "step_tflops_algorithmic": FLOPS_PER_TRIPLE * BATCH / (ms_d / d_steps * 1e-3) / 1e12}

    cpu = None
    if cpu_pool is not None:
        pool = cpu_pool
        try:
            v, total, wall = pool.run(64)
            tr_cpu, tr_cpu_dense = pool.run_tr(2)
        finally:
            pool.close()
        cores = pool.cores
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{total} instances of the same workload ({total // cores} per process x {cores} processes, 1 BLAS thread each, "
                         f"{wall:.1f} s wall)",
               "tr_iters_per_sec_random_starts": tr_cpu, "tr_iters_per_sec_dense_hessian": tr_cpu_dense,
               "tr_note": "tr_iters_per_sec_random_starts: one TR iteration per instance from RANDOM starts (tCG leaves after one product), "
                          "matrix-free Hessian operator; the like-for-like TR comparison on kicked Trotter starts is `tr.cpu`. "
                          "tr_iters_per_sec_dense_hessian: baseline (i), the reference's own dense Hessian (stiefel.py:498-529, 11 880 "
                          "jvp(grad) columns per iteration at this size), extrapolated from 4 timed columns per process"}

    # `value`: device-resident throughput of the better of the two ways to drive the device path (both timed above over
    # args.steps steps); the other one stays visible in `device_resident`
    device_resident = {"blocking": {"value": value, "ms_per_step": ms_per_step, "gpu_launches": int(launches),
                                    "mode": "one otn_triple call per step, device arrays, host synchronisation at the end of each call"},
                       "streamed": {"value": value_streamed, "ms_per_step": ms_dev_stream / args.steps, "gpu_launches": int(launches_stream),
                                    "mode": "otn_triple_submit / otn_triple_wait with device arrays (no copies), one step in flight "
                                            "ahead, two rotating output sets"}}
    best = "streamed"      # the way a device-resident caller drives the path (no host synchronisation between steps); fixed choice
    value, ms_per_step, launches = (device_resident[best][k] for k in ("value", "ms_per_step", "gpu_launches"))
    device_resident["value_is"] = best
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": n_warm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(), "host_numa_binding": numa,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps,
                    "mode": "streamed: otn_triple_submit / otn_triple_wait, one step in flight ahead, two rotating sets of pinned "
                            "host arrays; every step's H2D and D2H inside the timed region",
                    "synchronous": {"value": e2e_sync_value, "ms_per_step": ms_e2e_sync / args.steps,
                                    "mode": "one blocking otn_triple call per step"}},
            "gpu_launches": int(launches), "device_resident": device_resident,
            "roofline": roofline, "dense_formulation": dense_leg, "cpu_baseline": cpu, "tr": tr, "final_gather_ms": gather_ms,
            "two_site_model": two_site, "latency": latency, "complex_isometries": cplx, "c4": c4, "c5": c5}
    line["e2e"]["host_ceiling"] = host_ceiling
    line["e2e"]["frac_of_host_ceiling"] = e2e_value / host_ceiling["u1_per_s_ceiling"]
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-tr", action="store_true", help="skip the TR iterations/s leg")
    ap.add_argument("--no-dense", action="store_true", help="skip the dense-formulation comparison leg")
    ap.add_argument("--no-local", action="store_true", help="skip the 2-site-model (factored evaluation) leg")
    ap.add_argument("--no-latency", action="store_true", help="skip the config-1 / config-2 single-instance latency leg")
    ap.add_argument("--no-c4", action="store_true", help="skip the config-4 sweep leg")
    ap.add_argument("--no-complex", action="store_true", help="skip the complex-isometry U1 leg")
    ap.add_argument("--no-c5", action="store_true", help="skip the config-5 row-sharded N=6 composition leg")
    ap.add_argument("--min-warm-s", type=float, default=0.0,
                    help="keep warming up until this many seconds under load have passed (default 0 = exactly max(--warmup, 3) steps)")
    ap.add_argument("--workload", default="c3", choices=["c3", "c4", "c5"],
                    help="c3 (default): the headline line with every leg; c4 / c5: only that configuration's leg as the line")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
