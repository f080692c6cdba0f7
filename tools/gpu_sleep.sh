#!/bin/bash
# idle-role wait back-off in the pair kernel: 128 ns (main) vs 512 / 2000 ns builds
for v in main sleep512 sleep2000 main sleep512 sleep2000; do
  if [ $v = main ]; then unset TB200_LIB; else export TB200_LIB=$PWD/build/lib_$v.so; fi
  for cfg in "32 3 1 192000 1" "64 3 1 96000 0"; do
    set -- $cfg
    python tools/pair_micro.py $1 $2 $3 $4 64 $5 f16 4 fused 2>&1 | grep "fused:" | cut -c1-70 | sed "s/^/$v /"
  done
  for voc in bigvgan hifigan; do
  python bench.py --vocoder $voc --no-cpu-baseline --no-config4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v $voc', d['ms_per_step'])"
  done
done
