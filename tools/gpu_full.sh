#!/bin/bash
# Full parity suite + smoke + the default bench line (N = 1) on one GPU box.
tag=${1:-f0}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_tests.log 2>&1; echo "exit $?" >> $out/${tag}_tests.log
tail -4 $out/${tag}_tests.log
python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1; echo "exit $?" >> $out/${tag}_smoke.log
tail -3 $out/${tag}_smoke.log
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench exit $?"
cat $out/${tag}_bench.json | cut -c1-3000
