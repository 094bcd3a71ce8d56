#!/bin/bash
# Ablation timings of attn64_tc_kernel (which part binds?): builds variant libraries with -DA2_ABLATE=<bits> and times
# dpt_op_attention on the ViT-L shape. usage (GPU box): bash tools/attn_ablate.sh "0 1 2 3 4 8 16 31" [extra nvcc flags]
cd "$(dirname "$0")/.."
for a in ${1:-0 1 2 3}; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -I include -I muggled_dpt_b200/csrc \
    -DA2_ABLATE=$a $2 -o muggled_dpt_b200/lib/libdpt_b200_abl$a.so muggled_dpt_b200/csrc/dpt_api.cu || exit 1
  DPT_LIB=muggled_dpt_b200/lib/libdpt_b200_abl$a.so python tools/time_attention.py "ablate=$a"
done
