#!/bin/bash
# GPU box: the U1 legs of bench.py with the default kernels and with each A/B environment switch of libotn (one JSON line per run
# under gpurun_out/ab_<name>.json; summary table on stdout).
set -u
mkdir -p gpurun_out
F="--steps 20 --warmup 5 --no-cpu --no-dense --no-local --no-tr --no-latency --no-c4 --no-c5 --no-complex"
run() { name=$1; shift; env "$@" timeout 150 python bench.py $F > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err; }
run default OTN_AB=0
run back_v3 OTN_BACK_V3=1
run back_per_layer OTN_BACK_PER_LAYER=1
run chain_w8 OTN_CHAIN_W8=1
run asm_v1 OTN_ASM_V1=1
run tiled_chain OTN_FUSED_CHAIN=0
python - <<PY
import json
for n in ("default", "back_v3", "back_per_layer", "chain_w8", "asm_v1", "tiled_chain"):
    try:
        d = json.loads(open("gpurun_out/ab_%s.json" % n).read().strip().splitlines()[-1])
        r = d["roofline"]; o = r["other_kernels_ms_per_step"]
        print("%-15s value %7.0f  ms/step %.3f  e2e %7.0f  chain ms/step %.3f  asm %.3f  back %.3f  riem %.3f" % (
            n, d["value"], d["ms_per_step"], d["e2e"]["value"], r["kernel_share_of_step"] * r["profiled_ms_per_step"], o["assembly"], o["back_assembly"], o["riemannian_epilogue"]))
    except Exception as ex:
        print(n, "ERR", ex)
PY
