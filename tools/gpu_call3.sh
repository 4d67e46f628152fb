#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_train_gpu.py -q > gpurun_out/r2_pytest_train.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_train.txt
tail -40 gpurun_out/r2_pytest_train.txt
timeout 300 python tools/demo_loop.py > gpurun_out/r2_demo_loop.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_demo_loop.txt
tail -15 gpurun_out/r2_demo_loop.txt
