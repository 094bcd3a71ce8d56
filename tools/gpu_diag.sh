#!/bin/bash
# Runs every GPU test id in its own process (a trapped kernel poisons the CUDA context of its process only),
# each under a timeout, and writes a report to gpurun_out/diag.txt.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/diag.txt
: > $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv >> $OUT 2>&1
FILES="${@:-tests/test_ops_gpu.py tests/test_model_gpu.py}"
IDS=$(python -m pytest $FILES --collect-only -q -m gpu 2>/dev/null | grep '::' | sed 's/\[.*//' | awk '!seen[$0]++')
for id in $IDS; do
  echo "=== $id" >> $OUT
  timeout 300 python -m pytest "$id" -q -s -m gpu 2>&1 | grep -vE '^$|^=+ .* =+$|^platform|^rootdir|^plugins|^collected' | tail -40 >> $OUT
  echo "--- exit ${PIPESTATUS[0]}" >> $OUT
done
grep -cE '^--- exit 0' $OUT | xargs echo "passed:" | tee -a $OUT
grep -B30 -E '^--- exit [1-9]' $OUT | grep -E '^===|Error|error|assert|timeout' | head -80
tail -3 $OUT
