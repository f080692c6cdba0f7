#!/bin/bash
# Conv micro-benchmarks over the tuning knobs of conv1d_umma.cu (env: TB200_MAX_S, TB200_NPROD_SNAKE, TB200_NPROD_PW,
# TB200_L2_PREFETCH, TB200_TRACE).  usage: tools/tune_conv.sh <tag>   -> gpurun_out/<tag>.log
out=gpurun_out/${1:-tune}.log
: > $out
run() {
  python tools/conv_micro.py 32 32 3 1 0 192000 64 1 f16 3 >> $out 2>&1     # HiFiGAN stage 3
  python tools/conv_micro.py 64 64 11 1 0 96000 64 1 f16 3 >> $out 2>&1     # HiFiGAN stage 2
  python tools/conv_micro.py 32 32 3 1 0 192000 64 2 f16 3 >> $out 2>&1     # BigVGAN stage 3
  python tools/conv_micro.py 64 64 11 1 0 96000 64 2 f16 3 >> $out 2>&1     # BigVGAN stage 2
  python tools/conv_micro.py 128 128 7 1 0 24000 64 2 f16 3 >> $out 2>&1    # BigVGAN stage 1
  python tools/conv_micro.py 256 256 7 1 0 4000 64 2 f16 3 >> $out 2>&1     # BigVGAN stage 0
  python tools/conv_micro.py 192 1536 1 1 0 866 128 0 tf32 3 >> $out 2>&1   # acoustic FFN w1
}
echo "== default" >> $out; run
echo "== S<=4" >> $out; TB200_MAX_S=4 run
