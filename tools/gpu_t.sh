#!/bin/bash
# run a pytest selection on the GPU box: bash tools/gpu_t.sh <pytest args>
python -m pytest "$@" -x -q 2>&1 | tail -8
