#!/bin/bash
# ncu --set full of the four encoder GEMM launches of block 10 (qkv, proj, fc1, fc2) inside a ViT-L B=32 forward,
# then the sustained bench. usage: ncu_enc_gemms.sh <tag>
cd "$(dirname "$0")/.."
TAG=${1:-enc}
mkdir -p gpurun_out/$TAG
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py --no-cpu-baseline --dump-profile gpurun_out/$TAG/prof.csv > gpurun_out/$TAG/bench.json 2> gpurun_out/$TAG/bench.err
python -c "
import json; d=json.loads(open('gpurun_out/$TAG/bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']); print({k: round(v['ms_per_step'],3) for k,v in d['kernel_breakdown'].items()})"
G=$(grep -c "^gemm" gpurun_out/$TAG/prof.csv)   # gemm launches per step
FIRST=$(grep "^gemm" gpurun_out/$TAG/prof.csv | grep -n "blk10.qkv" | cut -d: -f1)
SKIP=$((3 * G + FIRST - 1))
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc --launch-skip $SKIP -c 4 -f -o gpurun_out/$TAG/ncu_gemm \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --profile-steps 0 > gpurun_out/$TAG/ncu_gemm.log 2>&1
ncu -i gpurun_out/$TAG/ncu_gemm.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
h=rows[0]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__issue_active.avg.pct']
idx=[(w,h.index(w)) for w in want if w in h]
for r in rows[2:]:
    print(' | '.join(f'{r[i][:60]}' for w,i in idx))
"
