#!/bin/bash
# A/B two libraries on the same box: bash tools/ab_bench.sh <lib_a> <lib_b>   ("main" = the in-tree library)
for v in $1 $2 $1 $2; do
  if [ $v = main ]; then unset TB200_LIB; else export TB200_LIB=build/lib_$v.so; fi
  python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', d['value'], d['ms_per_step'])"
done
unset TB200_LIB
TB200_TRACE=1 python tools/conv_micro.py 64 64 3 1 0 96000 64 2 f16 3 | sed -n '1,2p;7,9p' | cut -c1-700
python -m pytest tests/test_conv_gpu.py tests/test_vocoder_gpu.py -m gpu -x -q 2>&1 | tail -2
