#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_full.txt 2>&1; tail -5 gpurun_out/r2_pytest_full.txt
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; tail -3 gpurun_out/r2_bench_d.err
