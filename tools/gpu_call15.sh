#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_parity_gpu.py -q -x -s -k "full_size_opv2v" > gpurun_out/r2_pytest_i.txt 2>&1; tail -25 gpurun_out/r2_pytest_i.txt | cut -c1-220
