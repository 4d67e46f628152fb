#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_ops_gpu.py -q -x > gpurun_out/r2_pytest_g.txt 2>&1; tail -3 gpurun_out/r2_pytest_g.txt
timeout 600 python tools/exp_bw.py > gpurun_out/exp_bw5.txt 2>&1; grep "bn_stats\|MB per\|torch" gpurun_out/exp_bw5.txt
