#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_camera_gpu.py -q -x > gpurun_out/r2_pytest_h.txt 2>&1; tail -3 gpurun_out/r2_pytest_h.txt
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_train_step.csv python tools/prof_train.py 4 > gpurun_out/r2_prof_train.log 2>&1
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; tail -3 gpurun_out/r2_bench_d.err
