#!/bin/bash
# GPU box: stream tests only, then bench.py (no CPU / dense / 2-site / TR legs) with and without the environment switch $1
set -u
SW="${1:-OTN_NO_SPLIT=1}"
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_stream.py -x -q -m gpu > gpurun_out/t_abq.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_abq.log
timeout 100 python bench.py --no-cpu --no-dense --no-local --no-tr > gpurun_out/bq_default.json 2> gpurun_out/bq_default.err
env "$SW" timeout 100 python bench.py --no-cpu --no-dense --no-local --no-tr > gpurun_out/bq_switch.json 2> gpurun_out/bq_switch.err
python - <<PY
import json
for n in ("bq_default", "bq_switch"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % n).read().strip().splitlines()[-1])
        print(n, round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]))
    except Exception as ex:
        print(n, "ERR", ex, open("gpurun_out/%s.err" % n).read()[-1500:])
PY
