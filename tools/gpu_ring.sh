#!/bin/bash
mkdir -p gpurun_out
o=gpurun_out/r2_pair_ring.txt; : > $o
for mr in 0 6 10; do
  echo "TB200_PAIR_MIN_RING=$mr" >> $o
  for cfg in "64 7 1 96000 0" "64 11 1 96000 0" "32 7 1 192000 0" "128 3 1 48000 0" "64 7 1 96000 1" "32 7 1 192000 1" "128 3 1 48000 1"; do
    set -- $cfg
    TB200_PAIR_MIN_RING=$mr TB200_PLAN_DEBUG=1 python tools/pair_micro.py $1 $2 $3 $4 64 $5 f16 4 fused 2>&1 | grep -E "plan|fused:" | sort -u | cut -c1-150 >> $o
  done
done
cat $o
