#!/bin/bash
set -x
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -c 1800 gpurun_out/r2_bench_n2.json; tail -5 gpurun_out/r2_bench_n2.err
