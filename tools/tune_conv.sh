#!/bin/bash
# conv micro-benchmarks over tuning knobs
out=gpurun_out/${1:-tune}.log
: > $out
run() {
  python tools/conv_micro.py 32 32 3 1 0 192000 64 1 f16 3 >> $out 2>&1
  python tools/conv_micro.py 64 64 11 1 0 96000 64 1 f16 3 >> $out 2>&1
  python tools/conv_micro.py 32 32 3 1 0 192000 64 2 f16 3 >> $out 2>&1
  python tools/conv_micro.py 64 64 11 1 0 96000 64 2 f16 3 >> $out 2>&1
  python tools/conv_micro.py 128 128 7 1 0 24000 64 2 f16 3 >> $out 2>&1
}
echo "== default" >> $out; run
echo "== S<=4" >> $out; TB200_MAX_S=4 run
echo "== S<=2" >> $out; TB200_MAX_S=2 run
echo "== nprod snake 6" >> $out; TB200_NPROD_SNAKE=6 run
echo "== nprod pw 10" >> $out; TB200_NPROD_PW=10 run
