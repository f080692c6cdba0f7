#!/bin/bash
# N-GPU bench under torchrun: headline (weak scaling) + config-4 (one 512-utterance batch sharded by utterance).
# usage: bash tools/gpu_multi.sh N
n=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2_bench_n$n.json 2> gpurun_out/r2_bench_n$n.err
echo "exit $?"
tail -3 gpurun_out/r2_bench_n$n.err
cat gpurun_out/r2_bench_n$n.json
