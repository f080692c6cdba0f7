#!/bin/bash
mkdir -p gpurun_out
o=gpurun_out/r2_nprod_pw.txt; : > $o
for np in 0 10; do
  echo "TB200_NPROD_PW=$np" >> $o
  for sh in "1536 192 1 866" "384 1536 1 433" "192 384 5 433" "192 1536 1 866" "192 192 1 433"; do
    set -- $sh
    TB200_NPROD_PW=$np python tools/conv_micro.py $1 $2 $3 1 0 $4 128 0 tf32 5 2>&1 | grep TFLOP >> $o
  done
  TB200_NPROD_PW=$np python bench.py --workload acoustic --steps 5 --no-config4 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('acoustic ms', d['ms_per_step'])" >> $o
done
cat $o
