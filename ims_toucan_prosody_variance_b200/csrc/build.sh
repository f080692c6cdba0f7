#!/bin/bash
# Build libtoucan_b200.so for sm_100a (cross-compiles without a GPU).
set -e
here="$(cd "$(dirname "$0")" && pwd)"
out="${TB200_OUT:-$here/../libtoucan_b200.so}"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared \
     -Xptxas -v "$@" -o "$out" "$here/toucan_b200.cu" 2>&1
echo "built $out"
