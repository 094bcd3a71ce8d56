#!/bin/bash
# One GPU-box pass: parity tests, the default bench line, per-launch table, ncu launch list and ncu --set full captures
# of the top kernels. Outputs under gpurun_out/<tag>/ (tag = $1, default "run").
cd "$(dirname "$0")/.."
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
  timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.txt 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.txt
  tail -3 $OUT/pytest_gpu.txt
fi
timeout 600 python bench.py --dump-profile $OUT/launch_table_vitl_b32.csv > $OUT/bench_vitl_b32.json 2> $OUT/bench_vitl_b32.err
python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_vitl_b32.json").read().strip().splitlines()[-1])
    print("bench", round(d["value"], 1), "fps", round(d["ms_per_step"], 3), "ms e2e", d["e2e"] and round(d["e2e"]["value"], 1), "cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"])
    for k, v in sorted(d["kernel_breakdown"].items(), key=lambda kv: -kv[1]["ms_per_step"]):
        print(f"  {k:18s} {v['ms_per_step']:8.3f} ms  {v['tflops']:8.1f} TF/s {v['gbs']:8.0f} GB/s  x{v['launches_per_step']}")
except Exception as e:
    print("bench FAILED", e)
PY
if [ -z "$SKIP_CONFIGS" ]; then
  timeout 300 python bench.py --model vitb --batch 8 --no-cpu-baseline --dump-profile $OUT/launch_table_vitb.csv > $OUT/bench_vitb.json 2>> $OUT/bench_configs.err
  timeout 300 python bench.py --model beit_large_384 --batch 16 --size 384 --no-cpu-baseline --dump-profile $OUT/launch_table_beit_large_384.csv > $OUT/bench_beit_large_384.json 2>> $OUT/bench_configs.err
  timeout 300 python bench.py --model swinv2_large_384 --batch 16 --size 384 --dtype fp16 --no-cpu-baseline --dump-profile $OUT/launch_table_swinv2_large_384.csv > $OUT/bench_swinv2_large_384.json 2>> $OUT/bench_configs.err
  timeout 300 python bench.py --batch 4 --no-cpu-baseline > $OUT/bench_vitl_b4.json 2>> $OUT/bench_configs.err
  timeout 300 python tools/bench_prepost.py > $OUT/bench_prepost.json 2>> $OUT/bench_configs.err
  for f in vitb beit_large_384 swinv2_large_384 vitl_b4; do python -c "
import json; d=json.loads(open('$OUT/bench_$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value'],1), 'fps e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'])" 2>&1 | tail -1; done
  cat $OUT/bench_prepost.json
fi
if [ -z "$SKIP_NCU" ]; then
  # launch list of one timed step: 3 warm-up steps x (launches per step + the L2 flush) skipped
  LPS=$(( $(grep -c . $OUT/launch_table_vitl_b32.csv) ))   # table rows + header = launches + flush
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip $((3 * LPS)) -c $((2 * LPS)) --csv --log-file $OUT/ncu_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --profile-steps 0 > $OUT/ncu_launches.log 2>&1
  # full captures: one attention launch, the four encoder GEMM shapes, LayerNorm, the head kernels
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_tc --launch-skip 30 -c 1 -f -o $OUT/ncu_attn \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --profile-steps 0 > $OUT/ncu_attn.log 2>&1
  G=$(grep -c "^gemm" $OUT/launch_table_vitl_b32.csv)
  FIRST=$(grep "^gemm" $OUT/launch_table_vitl_b32.csv | grep -n "blk10.qkv" | cut -d: -f1)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc --launch-skip $((3 * G + FIRST - 1)) -c 4 -f -o $OUT/ncu_gemm \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --profile-steps 0 > $OUT/ncu_gemm.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"layernorm|resize|row_stats" --launch-skip 30 -c 10 -f -o $OUT/ncu_misc \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --profile-steps 0 > $OUT/ncu_misc.log 2>&1
  # keep the box output small (gpurun merges <= 64 MiB back): export the pages that get read, drop the reports
  for r in attn gemm misc; do
    ncu -i $OUT/ncu_$r.ncu-rep --page raw --csv > $OUT/ncu_${r}_raw.csv 2>/dev/null
  done
  ncu -i $OUT/ncu_attn.ncu-rep --page source --csv > $OUT/ncu_attn_source.csv 2>/dev/null
  rm -f $OUT/ncu_gemm.ncu-rep $OUT/ncu_misc.ncu-rep
fi
ls -la $OUT
