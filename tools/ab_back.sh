#!/bin/bash
# GPU box: parity tests that exercise the pull-back, then bench.py (U1 legs only) with the default kernel and with the switch $1
set -u
SW="${1:-OTN_BACK_V3=1}"
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_fused_chain.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py tests/test_gpu_tr.py tests/test_gpu_edge.py -x -q -m gpu > gpurun_out/t_ab.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/t_ab.log
F="--no-cpu --no-dense --no-local --no-tr --no-latency --no-c4 --no-c5 --no-complex"
timeout 120 python bench.py $F > gpurun_out/ab_default.json 2> gpurun_out/ab_default.err
env "$SW" timeout 120 python bench.py $F > gpurun_out/ab_switch.json 2> gpurun_out/ab_switch.err
python - <<PY
import json
for n in ("ab_default", "ab_switch"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % n).read().strip().splitlines()[-1])
        print(n, round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["roofline"].get("gemm_ms_per_step"), d["roofline"].get("other_kernels_ms_per_step"))
    except Exception as ex:
        print(n, "ERR", ex, open("gpurun_out/%s.err" % n).read()[-1500:])
PY
