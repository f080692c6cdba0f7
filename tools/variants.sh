#!/bin/bash
# Compare library variants (build/lib_<name>.so) on snake conv shapes, with the pipeline trace of the first one.
for v in "$@"; do
  echo "== $v"
  TB200_LIB=build/lib_$v.so TB200_TRACE=1 python tools/conv_micro.py 64 64 3 1 0 96000 64 2 f16 3 2>&1 | sed -n '1p;2p;7,9p' | cut -c1-1500
  TB200_LIB=build/lib_$v.so python tools/conv_micro.py 32 32 7 1 0 192000 64 2 f16 3 2>&1 | head -1
  TB200_LIB=build/lib_$v.so python tools/conv_micro.py 128 128 3 1 0 24000 64 2 f16 3 2>&1 | head -1
done
