#!/bin/bash
out=gpurun_out/${1:-tune}.log
: > $out
P=$PWD/ims_toucan_prosody_variance_b200
run() {
  python tools/conv_micro.py 64 64 11 1 0 96000 64 1 f16 3 >> $out 2>&1
  python tools/conv_micro.py 32 32 3 1 0 192000 64 1 f16 3 >> $out 2>&1
  python tools/conv_micro.py 128 128 7 1 0 24000 64 1 f16 3 >> $out 2>&1
  python tools/profile_vocoder.py hifigan 2>&1 | tail -11 | head -1 >> $out
}
echo "== pw 16 warps (6+8)" >> $out; run
echo "== pw 24 warps (14+8)" >> $out; TB200_LIB=$P/libtb_pw22.so run
