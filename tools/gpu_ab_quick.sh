#!/bin/bash
# tests of the conv paths + the four step times (BigVGAN, HiFiGAN, acoustic tf32 / f16)
python -m pytest tests/test_conv_gpu.py tests/test_vocoder_gpu.py tests/test_toucantts_gpu.py -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline --no-config4 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bigvgan ms', d['ms_per_step'])"
python bench.py --vocoder hifigan --no-cpu-baseline --no-config4 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('hifigan ms', d['ms_per_step'])"
for pr in tf32 f16; do
python bench.py --workload acoustic --steps 5 --no-config4 --no-cpu-baseline --acoustic-precision $pr 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('acoustic $pr ms', d['ms_per_step'])"
done
