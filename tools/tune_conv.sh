#!/bin/bash
out=gpurun_out/${1:-tune}.log
: > $out
python tools/conv_micro.py 32 32 3 1 0 192000 64 1 f16 3 >> $out 2>&1
python tools/conv_micro.py 64 64 11 1 0 96000 64 1 f16 3 >> $out 2>&1
python tools/conv_micro.py 64 64 11 1 0 96000 64 2 f16 3 >> $out 2>&1
python tools/conv_micro.py 32 32 3 1 0 192000 64 2 f16 3 >> $out 2>&1
python tools/conv_micro.py 128 128 3 1 0 24000 64 1 f16 3 >> $out 2>&1
