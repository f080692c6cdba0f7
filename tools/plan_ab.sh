#!/bin/bash
# A/B of the snake tiling rule (TB200_SNAKE_PLAN_OLD=1 disables it) on the same box.
for v in new old new old; do
  if [ $v = old ]; then export TB200_SNAKE_PLAN_OLD=1; else unset TB200_SNAKE_PLAN_OLD; fi
  python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', d['value'], d['ms_per_step'])"
done
unset TB200_SNAKE_PLAN_OLD
python -m pytest tests/test_conv_gpu.py tests/test_vocoder_gpu.py -m gpu -x -q 2>&1 | tail -2
