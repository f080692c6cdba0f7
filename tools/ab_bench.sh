#!/bin/bash
# A/B two libraries on the same box: bash tools/ab_bench.sh <lib_a> <lib_b>   ("main" = the in-tree library)
for wl in "--vocoder bigvgan" "--vocoder hifigan" "--workload acoustic --steps 5"; do
for v in $1 $2 $1 $2; do
  if [ $v = main ]; then unset TB200_LIB; else export TB200_LIB=build/lib_$v.so; fi
  python bench.py $wl --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl $v', d['value'], d['ms_per_step'])"
done
done
if [ $2 = main ]; then unset TB200_LIB; else export TB200_LIB=build/lib_$2.so; fi
python -m pytest tests/test_conv_gpu.py tests/test_vocoder_gpu.py tests/test_toucantts_gpu.py -m gpu -x -q 2>&1 | tail -2
