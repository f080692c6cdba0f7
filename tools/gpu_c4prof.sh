#!/bin/bash
mkdir -p gpurun_out
o=gpurun_out/r2_c4_profile_graphs.txt; : > $o
for n in 512 64; do for g in eager graphs; do
python tools/profile_c4.py $n 128 bigvgan $g >> $o 2>&1
done; done
cat $o
