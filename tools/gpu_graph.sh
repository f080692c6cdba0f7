#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_toucantts_gpu.py -k "cuda_graph" -x -q 2>&1 | tail -30 > gpurun_out/r2_graph_test.log
cat gpurun_out/r2_graph_test.log
