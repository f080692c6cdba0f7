#!/bin/bash
# One GPU-box pass: tests, bench, per-launch table, ncu launch list + full capture of the dominant conv shapes.
# usage: tools/gpu_round.sh <tag>
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_tests.log 2>&1; echo "exit $?" >> $out/${tag}_tests.log
python bench.py > $out/${tag}_bench.log 2>$out/${tag}_bench.err; echo "exit $?" >> $out/${tag}_bench.log
python bench.py --vocoder hifigan --no-cpu-baseline > $out/${tag}_bench_hifigan.log 2>&1
python tools/profile_vocoder.py bigvgan > $out/${tag}_prof_bigvgan.log 2>&1
python tools/profile_vocoder.py hifigan > $out/${tag}_prof_hifigan.log 2>&1
python bench.py --steps 1 --no-cpu-baseline > $out/${tag}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 234 -c 160 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 1 --no-cpu-baseline > $out/${tag}_ncu_launch.log 2>&1
M1="python tools/conv_micro.py 32 32 3 1 0 192000 64 2 f16 3"
$M1 > $out/${tag}_micro1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv1d_umma -s 3 -c 1 -f -o $out/${tag}_snake_c32k3 $M1 > $out/${tag}_ncu1.log 2>&1
M2="python tools/conv_micro.py 64 64 11 1 0 96000 64 1 f16 3"
$M2 > $out/${tag}_micro2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv1d_umma -s 3 -c 1 -f -o $out/${tag}_leaky_c64k11 $M2 > $out/${tag}_ncu2.log 2>&1
true
