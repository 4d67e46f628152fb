#!/bin/bash
# gpurun helper: the complete GPU test suite + the default bench line (what the driver runs at round end)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_full.txt 2>&1; tail -5 gpurun_out/r2_pytest_full.txt
timeout 900 python bench.py > gpurun_out/r2_bench_check.json 2> gpurun_out/r2_bench_check.err; tail -3 gpurun_out/r2_bench_check.err
