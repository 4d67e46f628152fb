#!/bin/bash
set -x
cd /root/repo 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_ops_gpu.py tests/test_train_gpu.py -q -x > gpurun_out/r2_pytest_g.txt 2>&1; tail -3 gpurun_out/r2_pytest_g.txt
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; tail -3 gpurun_out/r2_bench_d.err
