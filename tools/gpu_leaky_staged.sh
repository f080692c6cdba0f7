#!/bin/bash
# leaky single-conv staged mode vs the per-layer kernel, all pair shapes (dil 1)
mkdir -p gpurun_out
SNAKES="0" MODES="unfused staged" KDS="3,1 7,1 11,1" bash tools/pair_sweep.sh f16 > gpurun_out/r2_leaky_staged_sweep.txt 2>&1
cat gpurun_out/r2_leaky_staged_sweep.txt
