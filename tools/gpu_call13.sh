#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_camera_gpu.py -q > gpurun_out/r2_pytest_camera.txt 2>&1; grep -E "passed|failed|^FAILED|^E  " gpurun_out/r2_pytest_camera.txt | cut -c1-300
