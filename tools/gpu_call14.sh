#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -q -x > gpurun_out/r2_pytest_f.txt 2>&1; tail -3 gpurun_out/r2_pytest_f.txt
for cm in 0 1 2; do timeout 600 python tools/exp_conv_ops.py 12 $cm > gpurun_out/exp_conv_ops_cm$cm.txt 2>&1; tail -24 gpurun_out/exp_conv_ops_cm$cm.txt; done
timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; tail -3 gpurun_out/r2_bench_c.err
