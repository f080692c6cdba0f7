#!/bin/bash
# Build libtoucan_b200.so for sm_100a (cross-compiles without a GPU).
# Four translation units compiled in parallel, re-using objects whose sources have not changed:
#   toucan_b200.cu (C ABI + per-layer conv kernels + ragged kernels), acoustic.cu, respair_umma.cu, attention_umma.cu
set -e
here="$(cd "$(dirname "$0")" && pwd)"
out="${TB200_OUT:-$here/../libtoucan_b200.so}"
obj="${TB200_OBJ:-$here/../../build/obj}"
mkdir -p "$obj"
flags=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v "$@")
sig="$(echo "${flags[@]}" | md5sum | cut -c1-8)"
pids=()
for unit in toucan_b200 acoustic respair_umma attention_umma; do
  o="$obj/$unit.$sig.o"
  extra=()
  [ "$unit" != toucan_b200 ] && extra=(-DTB200_NO_AA_CONSTANT)
  if [ ! -f "$o" ] || [ -n "$(find "$here" "$here/../../include" -newer "$o" \( -name '*.cu' -o -name '*.cuh' -o -name '*.h' \) | head -1)" ]; then
    ( nvcc "${flags[@]}" "${extra[@]}" -c -o "$o.tmp" "$here/$unit.cu" > "$obj/$unit.log" 2>&1 && mv "$o.tmp" "$o" ) &
    pids+=($!)
  fi
done
rc=0
for p in "${pids[@]}"; do wait "$p" || rc=1; done
cat "$obj"/*.log 2>/dev/null
[ $rc -eq 0 ] || { echo "nvcc failed"; exit 1; }
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o "$out" "$obj/toucan_b200.$sig.o" "$obj/acoustic.$sig.o" "$obj/respair_umma.$sig.o" "$obj/attention_umma.$sig.o"
echo "built $out"
