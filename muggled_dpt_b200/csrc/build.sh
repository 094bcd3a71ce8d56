#!/bin/bash
# Builds libdpt_b200.so in-tree for sm_100a. Usage: build.sh [extra nvcc flags]
set -e
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
OUT="$HERE/../lib/libdpt_b200.so"
mkdir -p "$HERE/../lib"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  -Xcompiler -fPIC -shared -I "$ROOT/include" -I "$HERE" "$@" \
  -o "$OUT" "$HERE/dpt_api.cu"
echo "built $OUT"
