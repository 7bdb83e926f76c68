#!/bin/bash
# One GPU-box call: `ncu --set full` of the kernels of one block-diagonal U1 step (C3 workload), after the plain run exited 0.
set -u
mkdir -p gpurun_out
F="--steps 2 --warmup 3 --no-cpu --no-tr --no-dense --no-local --no-latency --no-c4 --no-c5 --no-complex"
timeout 300 python bench.py $F > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err; echo "plain rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:chain_fused|riem_stream|assemble_sym' -s 15 -c 5 \
    -o gpurun_out/step_full -f python bench.py $F > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/step_full.ncu-rep
