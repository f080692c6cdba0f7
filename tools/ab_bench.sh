for v in scalar pk3 scalar pk3; do
  TB200_LIB=build/lib_$v.so python bench.py --no-cpu-baseline > gpurun_out/ab_$v.log 2>/dev/null
  python -c "
import json
d=json.loads(open('gpurun_out/ab_$v.log').read().strip().splitlines()[-1]); print('$v', d['value'], d['ms_per_step'])"
done
TB200_LIB=build/lib_pk3.so python -m pytest tests/test_conv_gpu.py tests/test_vocoder_gpu.py -m gpu -x -q 2>&1 | tail -2
