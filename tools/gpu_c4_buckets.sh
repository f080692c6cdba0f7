#!/bin/bash
# config-4 bucket policy: utterances x (max_batch, padding ratio) on one GPU
mkdir -p gpurun_out
o=gpurun_out/r2_c4_buckets.txt; : > $o
for n in 64 512; do
  for mb_pr in "64 1.5" "64 3" "64 1000" "128 1000" "256 1000"; do
    set -- $mb_pr
    [ $n = 64 ] && [ $1 != 64 ] && continue
    echo "utterances=$n max_batch=$1 padding_ratio=$2" >> $o
    python bench.py --config4-only --config4-utterances $n --config4-max-batch $1 --config4-padding-ratio $2 2>&1 | tail -1 | cut -c1-420 >> $o
  done
done
cat $o
