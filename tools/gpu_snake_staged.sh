#!/bin/bash
# snake single-conv staged mode vs the per-layer kernel, C = 64 / 128 / 32 pair shapes
mkdir -p gpurun_out
SNAKES="1" MODES="unfused staged" CS="64 128 32" KDS="3,1 7,1 11,1 11,5" bash tools/pair_sweep.sh f16 > gpurun_out/r2_snake_staged_sweep.txt 2>&1
grep -v "^C=.*->" gpurun_out/r2_snake_staged_sweep.txt | cut -c1-120
