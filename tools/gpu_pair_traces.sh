#!/bin/bash
mkdir -p gpurun_out
o=gpurun_out/r2_pair_traces_final.txt; : > $o
for cfg in "64 7 1 96000 1" "64 7 1 96000 0" "32 3 1 192000 1" "64 3 1 96000 1"; do
  set -- $cfg
  TB200_TRACE=1 TB200_PLAN_DEBUG=1 python tools/pair_micro.py $1 $2 $3 $4 64 $5 f16 3 fused 2>&1 | grep -v "^$" | head -14 | cut -c1-230 >> $o
done
cat $o
