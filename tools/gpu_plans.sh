#!/bin/bash
# plans and timings of the acoustic model's main GEMM shapes (tf32 and f16 operands), full-length batches
mkdir -p gpurun_out
o=gpurun_out/r2_acoustic_gemm_plans.txt; : > $o
for prec in tf32 f16; do
for sh in "192 384 5 433" "192 192 1 433" "384 1536 1 433" "1536 192 1 866" "192 1536 1 866" "192 576 1 866" "256 256 5 866"; do
  set -- $sh
  TB200_PLAN_DEBUG=1 python tools/conv_micro.py $1 $2 $3 1 0 $4 128 0 $prec 5 2>&1 | grep -E "tb200 plan|TFLOP" | sort -u >> $o
done; done
cat $o
