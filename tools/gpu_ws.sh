#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_vocoder_gpu.py -k "workspace_garbage" -x -q 2>&1 | tail -30 > gpurun_out/r2_ws_test.log
python -m pytest tests/test_respair_gpu.py -k "beyond_32" -x -q 2>&1 | tail -15 >> gpurun_out/r2_ws_test.log
cat gpurun_out/r2_ws_test.log
