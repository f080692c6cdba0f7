#!/bin/bash
mkdir -p gpurun_out
o=gpurun_out/r2_trace_acoustic_gemms.txt; : > $o
for sh in "1536 192 1 866" "384 1536 1 433" "192 384 5 433"; do
  set -- $sh
  TB200_PLAN_DEBUG=1 TB200_TRACE=1 python tools/conv_micro.py $1 $2 $3 1 0 $4 128 0 tf32 3 2>&1 | head -16 | cut -c1-230 >> $o
done
cat $o
