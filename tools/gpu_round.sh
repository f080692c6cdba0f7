#!/bin/bash
# One GPU-box pass: tests, smoke, bench (all workloads), per-launch table, ncu launch list with DRAM bytes.
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_tests.log 2>&1; echo "exit $?" >> $out/${tag}_tests.log
python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1; echo "exit $?" >> $out/${tag}_smoke.log
python bench.py > $out/${tag}_bench.log 2>$out/${tag}_bench.err; echo "exit $?" >> $out/${tag}_bench.log
python bench.py --vocoder hifigan --no-cpu-baseline > $out/${tag}_bench_hifigan.log 2>&1
python bench.py --workload acoustic --steps 5 > $out/${tag}_bench_acoustic.log 2>&1
python bench.py --workload e2e --steps 5 > $out/${tag}_bench_e2e.log 2>&1
python bench.py --impl reference --steps 1 --warmup 0 > $out/${tag}_bench_ref.log 2>&1
python tools/profile_vocoder.py bigvgan > $out/${tag}_prof_bigvgan.log 2>&1
# launch list of one timed step (warm-up = pack kernels + 3 forwards; skip them), with DRAM bytes per launch
python bench.py --steps 1 --no-cpu-baseline > $out/${tag}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:conv1d_umma -s 234 -c 78 \
    --csv --log-file $out/${tag}_launches.csv python bench.py --steps 1 --no-cpu-baseline > $out/${tag}_ncu_launch.log 2>&1
true
