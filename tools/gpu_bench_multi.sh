#!/bin/bash
set -x
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -c 600 gpurun_out/r2_bench_n$N.json; tail -5 gpurun_out/r2_bench_n$N.err
