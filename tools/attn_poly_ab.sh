#!/bin/bash
# A/B of the attention kernel's FMA-pipe exp2 share: muggled_dpt_b200/lib/libdpt_poly<P>.so built with -DATT_POLY_PAIRS=P
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/poly
python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "attention" 2>&1 | tail -3
for P in ${VARIANTS:-0 2 3 4}; do
  for B in ${BATCHES:-32}; do
    DPT_B200_LIB=$PWD/muggled_dpt_b200/lib/libdpt_poly$P.so timeout 300 python bench.py --batch $B --steps 8 --warmup 3 --no-cpu-baseline --no-e2e \
      --dump-profile gpurun_out/poly/prof_p${P}_b$B.csv > gpurun_out/poly/bench_p${P}_b$B.json 2> gpurun_out/poly/err_p${P}_b$B.txt
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/poly/bench_p${P}_b$B.json").read().strip().splitlines()[-1])
    print("poly", $P, "B", $B, round(d["value"], 1), "fps", "attn ms/step", round(d["kernel_breakdown"]["attn"]["ms_per_step"], 3), "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("poly", $P, "FAILED", e)
PY
  done
done
