#!/bin/bash
# Compare library variants (build/lib_<name>.so) on representative conv shapes.
for v in "$@"; do
  echo "== $v"
  for shape in "64 64 3 1 0 96000 64 2" "64 64 11 1 0 96000 64 2" "128 128 11 1 0 24000 64 2" "256 256 7 1 0 4000 64 2" "32 32 11 1 0 192000 64 2" "64 64 11 1 0 96000 64 1" "128 64 8 1 4 24000 64 0"; do
    TB200_LIB=build/lib_$v.so python tools/conv_micro.py $shape f16 3 2>&1 | head -1
  done
done
