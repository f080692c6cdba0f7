#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_conv_gpu.py tests/test_toucantts_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -4
for pr in tf32 f16; do
python bench.py --workload acoustic --steps 5 --no-config4 --no-cpu-baseline --acoustic-precision $pr 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('acoustic $pr ms', d['ms_per_step'])"
done
