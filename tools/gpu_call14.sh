#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_full.txt 2>&1; tail -4 gpurun_out/r2_pytest_full.txt
timeout 600 python tools/exp_conv_ops.py 12 > gpurun_out/exp_conv_ops2.txt 2>&1; cat gpurun_out/exp_conv_ops2.txt | tail -24
timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; tail -3 gpurun_out/r2_bench_b.err
