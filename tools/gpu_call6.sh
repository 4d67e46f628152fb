#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_ops_gpu.py tests/test_train_gpu.py -q > gpurun_out/r2_pytest_train.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_train.txt
grep -E "passed|failed|^FAILED|^E  " gpurun_out/r2_pytest_train.txt | cut -c1-250
timeout 600 python -m pytest tests/test_parity_gpu.py -q -k "warp_fuse_op_golden or bf16_rounded" > gpurun_out/r2_pytest_b.txt 2>&1; grep -E "passed|failed|^FAILED|^E  " gpurun_out/r2_pytest_b.txt | cut -c1-250
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_train_step.csv python tools/prof_train.py 4 > gpurun_out/r2_prof_train.log 2>&1
timeout 900 python bench.py --steps 30 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_train_n1.json 2> gpurun_out/r2_bench_train_n1.err
tail -c 900 gpurun_out/r2_bench_train_n1.json; tail -5 gpurun_out/r2_bench_train_n1.err
