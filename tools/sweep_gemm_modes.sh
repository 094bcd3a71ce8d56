#!/bin/bash
# Per-launch GEMM timings for each tile mode (DPT_GEMM_MODE, dpt_api.cu) over a range of batch sizes.
# Writes gpurun_out/sweep/prof_m<mode>_b<B>.csv and bench_m<mode>_b<B>.json; tools/fit_gemm_modes.py summarises them.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/sweep
MODEL=${MODEL:-vitl}
for B in ${BATCHES:-1 2 4 8 16 32}; do
  for M in ${MODES:-0 1 2 3}; do
    DPT_GEMM_MODE=$M timeout 300 python bench.py --model $MODEL --batch $B --steps 8 --warmup 3 --no-cpu-baseline --no-e2e \
      --dump-profile gpurun_out/sweep/prof_${MODEL}_m${M}_b${B}.csv > gpurun_out/sweep/bench_${MODEL}_m${M}_b${B}.json 2> gpurun_out/sweep/err_${MODEL}_m${M}_b${B}.txt
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/sweep/bench_${MODEL}_m${M}_b${B}.json").read().strip().splitlines()[-1])
    print("$MODEL B", $B, "mode", $M, round(d["value"], 1), "fps", round(d["ms_per_step"], 3), "ms")
except Exception as e:
    print("$MODEL B", $B, "mode", $M, "FAILED", e)
PY
  done
done
