#!/bin/bash
# One GPU-box pass: tests, smoke, bench (all workloads), per-launch tables, ncu launch list with DRAM bytes,
# full ncu captures of two representative conv shapes, per-tile pipeline traces.
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_tests.log 2>&1; echo "exit $?" >> $out/${tag}_tests.log
python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1; echo "exit $?" >> $out/${tag}_smoke.log
python bench.py > $out/${tag}_bench.log 2>$out/${tag}_bench.err; echo "exit $?" >> $out/${tag}_bench.log
python bench.py --vocoder hifigan --no-cpu-baseline > $out/${tag}_bench_hifigan.log 2>&1
python bench.py --workload acoustic --steps 5 > $out/${tag}_bench_acoustic.log 2>&1
python bench.py --workload acoustic --steps 5 --acoustic-precision f16 --no-cpu-baseline > $out/${tag}_bench_acoustic_f16.log 2>&1
python bench.py --workload e2e --steps 5 > $out/${tag}_bench_e2e.log 2>&1
python bench.py --impl reference --steps 1 --warmup 0 > $out/${tag}_bench_ref.log 2>&1
python tools/profile_vocoder.py bigvgan > $out/${tag}_prof_bigvgan.log 2>&1
python tools/profile_vocoder.py hifigan > $out/${tag}_prof_hifigan.log 2>&1
TB200_TRACE=1 python tools/conv_micro.py 64 64 3 1 0 96000 64 2 f16 1 > $out/${tag}_trace_snake_c64k3.log 2>&1
TB200_TRACE=1 python tools/conv_micro.py 64 64 3 1 0 96000 64 1 f16 1 > $out/${tag}_trace_leaky_c64k3.log 2>&1
python bench.py --steps 1 --no-cpu-baseline > $out/${tag}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:conv1d_umma -s 234 -c 78 \
    --csv --log-file $out/${tag}_launches.csv python bench.py --steps 1 --no-cpu-baseline > $out/${tag}_ncu_launch.log 2>&1
M1="python tools/conv_micro.py 32 32 3 1 0 192000 64 1 f16 3"
$M1 > $out/${tag}_micro1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv1d_umma -s 3 -c 1 -f -o $out/${tag}_leaky_c32k3 $M1 > $out/${tag}_ncu1.log 2>&1
M2="python tools/conv_micro.py 64 64 11 1 0 96000 64 2 f16 3"
$M2 > $out/${tag}_micro2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv1d_umma -s 3 -c 1 -f -o $out/${tag}_snake_c64k11 $M2 > $out/${tag}_ncu2.log 2>&1
true
