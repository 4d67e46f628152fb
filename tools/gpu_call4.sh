#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -q > gpurun_out/r2_pytest_train.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_train.txt
grep -E "passed|failed" gpurun_out/r2_pytest_train.txt
timeout 900 python bench.py --steps 30 --warmup 5 --no-extras > gpurun_out/r2_bench_train_n1.json 2> gpurun_out/r2_bench_train_n1.err
tail -c 1500 gpurun_out/r2_bench_train_n1.json; tail -20 gpurun_out/r2_bench_train_n1.err
