#!/bin/bash
# ncu --set full of the decoder launches (reassembly, fusion, head incl. the halo convolution) of one ViT-L B=32 forward.
cd "$(dirname "$0")/.."
TAG=${1:-dec}
mkdir -p gpurun_out/$TAG
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --dump-profile gpurun_out/$TAG/prof.csv > gpurun_out/$TAG/bench.json 2> gpurun_out/$TAG/bench.err
N=$(grep -cE "^(gemm|conv_halo)" gpurun_out/$TAG/prof.csv)
FIRST=$(grep -E "^(gemm|conv_halo)" gpurun_out/$TAG/prof.csv | grep -n "reasm0.proj1x1" | cut -d: -f1)
timeout 900 ncu --set full --clock-control none -k regex:"gemm_tc|conv3x3_halo" --launch-skip $((3 * N + FIRST - 1)) -c $((N - FIRST + 1)) -f -o gpurun_out/$TAG/ncu_dec \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --profile-steps 0 > gpurun_out/$TAG/ncu_dec.log 2>&1
ncu -i gpurun_out/$TAG/ncu_dec.ncu-rep --page raw --csv > gpurun_out/$TAG/ncu_dec_raw.csv 2>/dev/null
rm -f gpurun_out/$TAG/ncu_dec.ncu-rep
grep -E "^(gemm|conv_halo)" gpurun_out/$TAG/prof.csv | tail -n +$FIRST | cut -d, -f1 > gpurun_out/$TAG/dec_labels.txt
wc -l gpurun_out/$TAG/dec_labels.txt gpurun_out/$TAG/ncu_dec_raw.csv
