#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_camera_gpu.py -q > gpurun_out/r2_pytest_camera.txt 2>&1; grep -E "passed|failed|^FAILED|^E  " gpurun_out/r2_pytest_camera.txt | cut -c1-300
timeout 600 python -m pytest tests/test_parity_gpu.py -q -k "precise_mode_matches or bf16_mode_tensor or conv_tile or channel_major or full_size_opv2v_two" > gpurun_out/r2_pytest_d.txt 2>&1; grep -E "passed|failed|^FAILED|^E  " gpurun_out/r2_pytest_d.txt | cut -c1-300
