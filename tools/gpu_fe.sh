#!/bin/bash
python -m pytest tests/test_pipeline_gpu.py -x -q 2>&1 | tail -5
