#!/bin/bash
# Short GPU-box pass: conv + vocoder parity tests, both vocoder benches, one pipeline trace.
tag=${1:-qX}
out=gpurun_out
mkdir -p $out
python -m pytest tests/test_conv_gpu.py tests/test_vocoder_gpu.py -m gpu -x -q > $out/${tag}_tests.log 2>&1; echo "exit $?" >> $out/${tag}_tests.log
tail -3 $out/${tag}_tests.log
python bench.py --no-cpu-baseline > $out/${tag}_bench.log 2>$out/${tag}_bench.err
python bench.py --vocoder hifigan --no-cpu-baseline > $out/${tag}_bench_hifigan.log 2>&1
python -c "
import json,sys
for f in ['$out/${tag}_bench.log','$out/${tag}_bench_hifigan.log']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
"
TB200_TRACE=1 python tools/conv_micro.py 64 64 3 1 0 96000 64 2 f16 1 > $out/${tag}_trace_snake_c64k3.log 2>&1
head -8 $out/${tag}_trace_snake_c64k3.log | cut -c1-220
