#!/bin/bash
# Round-2 GPU pass (one gpurun call): bench lines of all BASELINE configs, the reference arms, per-launch tables, the ncu
# launch list of the headline config and ncu --set full captures of the attention / GEMM / memory-bound kernels of the
# ViT-L, BEiT-L and SwinV2-L paths, the five-config parity report. Outputs under gpurun_out/<tag>/ (tag = $1).
# ncu runs use DPT_GRAPH=0 so that --launch-skip counts plain launches.
cd "$(dirname "$0")/.."
TAG=${1:-r2final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
summ() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1]); e=d.get('e2e') or {}
print('$2', round(d['value'],1), 'fps e2e', round(e.get('value',0),1), (d.get('clocks') or {}).get('sm_mhz'))" 2>&1 | tail -1; }
timeout 600 python bench.py --steps 20 --warmup 3 --dump-profile $OUT/launch_table_vitl_b32.csv > $OUT/bench_vitl_b32.json 2> $OUT/bench_vitl_b32.err; summ $OUT/bench_vitl_b32.json vitl_b32
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/bench_reference_cpu.json 2>> $OUT/ref.err; summ $OUT/bench_reference_cpu.json reference_cpu
timeout 600 python bench.py --impl reference-gpu --steps 10 --warmup 3 > $OUT/bench_reference_gpu_eager.json 2>> $OUT/ref.err; summ $OUT/bench_reference_gpu_eager.json reference_gpu_eager
timeout 300 python bench.py --model vitb --batch 8 --steps 20 --no-cpu-baseline --dump-profile $OUT/launch_table_vitb.csv > $OUT/bench_vitb.json 2>> $OUT/bench_configs.err; summ $OUT/bench_vitb.json vitb_b8
timeout 300 python bench.py --model beit_large_384 --batch 16 --size 384 --steps 20 --no-cpu-baseline --dump-profile $OUT/launch_table_beit_large_384.csv > $OUT/bench_beit_large_384.json 2>> $OUT/bench_configs.err; summ $OUT/bench_beit_large_384.json beit_large
timeout 300 python bench.py --model swinv2_large_384 --batch 16 --size 384 --dtype fp16 --steps 20 --no-cpu-baseline --dump-profile $OUT/launch_table_swinv2_large_384.csv > $OUT/bench_swinv2_large_384.json 2>> $OUT/bench_configs.err; summ $OUT/bench_swinv2_large_384.json swinv2_large
timeout 300 python bench.py --batch 4 --steps 20 --no-cpu-baseline --dump-profile $OUT/launch_table_vitl_b4.csv > $OUT/bench_vitl_b4.json 2>> $OUT/bench_configs.err; summ $OUT/bench_vitl_b4.json vitl_b4
timeout 300 python bench.py --batch 1 --steps 20 --warmup 5 --no-cpu-baseline --dump-profile $OUT/launch_table_vitl_b1.csv > $OUT/bench_vitl_b1.json 2>> $OUT/bench_configs.err; summ $OUT/bench_vitl_b1.json vitl_b1
timeout 300 python bench.py --model vits --batch 1 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_vits_b1.json 2>> $OUT/bench_configs.err; summ $OUT/bench_vits_b1.json vits_b1
timeout 300 python bench.py --model vits --batch 32 --steps 20 --no-cpu-baseline > $OUT/bench_vits_b32.json 2>> $OUT/bench_configs.err; summ $OUT/bench_vits_b32.json vits_b32
timeout 400 python bench.py --model vitg --batch 16 --steps 5 --warmup 3 --no-cpu-baseline --dump-profile $OUT/launch_table_vitg_b16.csv > $OUT/bench_vitg_b16.json 2>> $OUT/bench_configs.err; summ $OUT/bench_vitg_b16.json vitg_b16
timeout 300 python tools/bench_prepost.py > $OUT/bench_prepost.json 2>> $OUT/bench_configs.err
timeout 600 python tools/parity_report.py $OUT/parity_report.json SBLWE > $OUT/parity_report.txt 2>&1; tail -10 $OUT/parity_report.txt
export DPT_GRAPH=0
LPS=$(( $(grep -c . $OUT/launch_table_vitl_b32.csv) ))   # table rows + header = launches + flush
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip $((3 * LPS + 2)) -c $((2 * LPS)) --csv --log-file $OUT/ncu_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --profile-steps 0 > $OUT/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn64 --launch-skip 30 -c 1 -f -o $OUT/ncu_attn \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --profile-steps 0 > $OUT/ncu_attn.log 2>&1
G=$(grep -c "^gemm" $OUT/launch_table_vitl_b32.csv)
FIRST=$(grep "^gemm" $OUT/launch_table_vitl_b32.csv | grep -n "blk10.qkv" | cut -d: -f1)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc --launch-skip $((3 * G + FIRST - 1)) -c 4 -f -o $OUT/ncu_gemm \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --profile-steps 0 > $OUT/ncu_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"layernorm|resize|row_stats|conv3x3_halo" --launch-skip 33 -c 11 -f -o $OUT/ncu_misc \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --profile-steps 0 > $OUT/ncu_misc.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:attn64 --launch-skip 30 -c 1 -f -o $OUT/ncu_attn_beit \
  python bench.py --model beit_large_384 --batch 16 --size 384 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --profile-steps 0 > $OUT/ncu_attn_beit.log 2>&1
# SwinV2-L: one stage-2 block (block index >= 4): its seven launches (qkv GEMM with the q/k normalise epilogue, attention,
# proj, post-norm + fc1 operand, fc1, fc2, post-norm + next block's windowed operand)
timeout 900 ncu --set full --clock-control none -k regex:"attn64|swin_ln_residual|gemm_tc" --launch-skip $((3 * 180 + 60)) -c 7 -f -o $OUT/ncu_swin \
  python bench.py --model swinv2_large_384 --batch 16 --size 384 --dtype fp16 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --profile-steps 0 > $OUT/ncu_swin.log 2>&1
for r in attn gemm misc attn_beit swin; do
  ncu -i $OUT/ncu_$r.ncu-rep --page raw --csv > $OUT/ncu_${r}_raw.csv 2>/dev/null
done
ncu -i $OUT/ncu_attn.ncu-rep --page source --csv > $OUT/ncu_attn_source.csv 2>/dev/null
rm -f $OUT/*.ncu-rep
ls -la $OUT | head -50
