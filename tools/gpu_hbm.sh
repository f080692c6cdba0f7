#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_acoustic_kernels_gpu.py tests/test_toucantts_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_hbm_tests.log
cat gpurun_out/r2_hbm_tests.log
