#!/bin/bash
# GPU box: full GPU test-suite, then bench.py twice: default, and with the environment switch given as $1 (e.g. OTN_RIEM_STAGED=1)
set -u
SW="${1:-OTN_RIEM_STAGED=1}"
mkdir -p gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/t_ab.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/t_ab.log
timeout 150 python bench.py --no-cpu --no-dense --no-local > gpurun_out/b_default.json 2> gpurun_out/b_default.err
env "$SW" timeout 150 python bench.py --no-cpu --no-dense --no-local > gpurun_out/b_switch.json 2> gpurun_out/b_switch.err
python - <<PY
import json
for n in ("b_default", "b_switch"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % n).read().strip().splitlines()[-1])
        print(n, round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["roofline"]["other_kernels_ms_per_step"], round(d["tr"]["iters_per_sec"]))
    except Exception as ex:
        print(n, "ERR", ex, open("gpurun_out/%s.err" % n).read()[-1500:])
PY
