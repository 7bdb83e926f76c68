#!/bin/bash
# One GPU-box call: GPU test-suite, both bench arms with the driver's flags, the ncu launch list of two timed steps and ONE
# `ncu --set full` capture (argument: "step" = kernels of the U1 step, "fact" = the 2-site factored kernel; gpurun brings back at
# most 64 MiB, one report per call).  Numbers printed under ncu are not bench values.
set -u
WHAT="${1:-step}"
mkdir -p gpurun_out
if [ "$WHAT" = "step" ]; then
timeout 900 python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/t_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_final.log
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; echo "ref rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
F="--steps 2 --warmup 3 --no-cpu --no-tr --no-dense --no-local --no-latency --no-c4 --no-c5 --no-complex"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py $F > gpurun_out/ncu_final.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:chain_fused|riem_stream|assemble_sym' -s 15 -c 5 \
    -o gpurun_out/step_full_final -f python bench.py $F > gpurun_out/ncu_full_final.log 2>&1; echo "ncu full rc=$?"
timeout 200 python -c "from __graft_entry__ import smoke; smoke()" 2>&1 | tail -2
else
# the 2-site (factored) evaluation's kernel, for the record of configs 1 / 2 / 4
G="--steps 2 --warmup 3 --no-cpu --no-tr --no-dense --no-latency --no-c4 --no-c5 --no-complex"
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:fact_kernel' -c 2 \
    -o gpurun_out/fact_full_final -f python bench.py $G > gpurun_out/ncu_fact_final.log 2>&1; echo "ncu fact rc=$?"
fi
ls -la gpurun_out
