#!/bin/bash
python -m pytest tests/test_toucantts_gpu.py -k "flip_rate" -x -q -s 2>&1 | tail -12
