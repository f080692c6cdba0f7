for r in 0.3 0.05 0.3 0.05; do
  TB200_SNAKE_A2_RATIO=$r python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ratio $r', d['value'], d['ms_per_step'])"
done
