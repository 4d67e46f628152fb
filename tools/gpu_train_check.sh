#!/bin/bash
# gpurun helper: training op / step tests + smoke + one bench line with the train object
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_ops_gpu.py tests/test_train_gpu.py -q -x > gpurun_out/r2_pytest_g.txt 2>&1; tail -3 gpurun_out/r2_pytest_g.txt
timeout 600 python __graft_entry__.py --smoke > gpurun_out/r2_smoke.txt 2>&1; tail -1 gpurun_out/r2_smoke.txt
