#!/bin/bash
# One GPU-box call: GPU test-suite, both bench arms, then the ncu launch list of two timed steps (numbers under ncu are not bench values).
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/t_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_final.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; echo "ref rc=$?"
timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 2 --warmup 3 --min-warm-s 0 --no-cpu --no-tr --no-dense --no-local > gpurun_out/ncu_final.log 2>&1; echo "ncu rc=$?"
python -c "from __graft_entry__ import smoke; smoke()" 2>&1 | tail -2
