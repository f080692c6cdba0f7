#!/bin/bash
# Round-end GPU pass: parity suite, smoke, every bench workload, per-layer tables, ncu launch list of one timed step
# (after the same command exited 0 without ncu), full ncu captures of one fused pair and one per-layer snake conv.
tag=${1:-final}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_tests.log 2>&1; echo "exit $?" >> $out/${tag}_tests.log
tail -3 $out/${tag}_tests.log
python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1; echo "exit $?" >> $out/${tag}_smoke.log
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench exit $?"
python bench.py --vocoder hifigan --no-cpu-baseline --no-config4 > $out/${tag}_bench_hifigan.json 2>&1
python bench.py --workload acoustic --steps 5 --no-config4 > $out/${tag}_bench_acoustic.json 2>&1
python bench.py --workload acoustic --steps 5 --acoustic-precision tf32 --no-cpu-baseline --no-config4 > $out/${tag}_bench_acoustic_tf32.json 2>&1
python bench.py --workload e2e --steps 5 --no-config4 > $out/${tag}_bench_e2e.json 2>&1
python bench.py --impl reference --steps 1 --warmup 0 > $out/${tag}_bench_ref.json 2>&1
python tools/profile_vocoder.py bigvgan 64 500 f16 f16 > $out/${tag}_prof_bigvgan.log 2>&1
python tools/profile_vocoder.py hifigan 64 500 f16 f16 > $out/${tag}_prof_hifigan.log 2>&1
python tools/profile_tts.py 128 tf32 > $out/${tag}_prof_tts.log 2>&1
P="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-config4"
$P > $out/${tag}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"conv1d_umma|respair" -s 198 -c 66 \
    --csv --log-file $out/${tag}_launches.csv $P > $out/${tag}_ncu_launch.log 2>&1
M1="python tools/pair_micro.py 32 3 1 192000 64 1 f16 3 fused"
$M1 > $out/${tag}_micro1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:respair -s 3 -c 1 -f -o $out/${tag}_respair_snake_c32k3 $M1 > $out/${tag}_ncu1.log 2>&1
M2="python tools/conv_micro.py 64 64 7 1 0 96000 64 2 f16 3"
$M2 > $out/${tag}_micro2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv1d_umma -s 3 -c 1 -f -o $out/${tag}_conv_snake_c64k7 $M2 > $out/${tag}_ncu2.log 2>&1
ls -la $out/${tag}_* | head -40
true
