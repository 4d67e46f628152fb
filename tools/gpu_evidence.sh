#!/bin/bash
# final round-2 evidence: bench (default, precise, reference arm), launch lists, ncu --set full of the final conv / HBM kernels
# (exported to CSV on the box: .ncu-rep files are too large to travel back), sanitizer, smoke
set -x
mkdir -p gpurun_out
nproc > gpurun_out/r2_env.txt; nvidia-smi -L >> gpurun_out/r2_env.txt; nvidia-smi --query-gpu=power.limit,power.max_limit,clocks.max.sm --format=csv >> gpurun_out/r2_env.txt
timeout 600 python __graft_entry__.py --smoke > gpurun_out/r2_smoke.txt 2>&1; tail -2 gpurun_out/r2_smoke.txt
timeout 600 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
timeout 600 python bench.py --precise --steps 30 --warmup 5 --no-extras > gpurun_out/r2_bench_precise.json 2> gpurun_out/r2_bench_precise.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_cmd.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:conv_gemm -c 45 -f -o /tmp/r2_conv_full \
    python tools/prof_step.py 12 > gpurun_out/r2_ncu_conv.log 2>&1
ncu -i /tmp/r2_conv_full.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_conv_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:'warp_att|vox2|canvas_clear|pfn' -c 24 -f -o /tmp/r2_hbm_full \
    python tools/prof_step.py 12 > gpurun_out/r2_ncu_hbm.log 2>&1
ncu -i /tmp/r2_hbm_full.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_hbm_raw.csv 2>/dev/null
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py tests/test_train_ops_gpu.py -x -q \
    -k "precise_mode_matches_reference_golden or bf16_mode_tensor_core or varying_cloud or channel_major or batchnorm or wgrad_tensor_core" > gpurun_out/r2_sanitizer_memcheck.txt 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2_sanitizer_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py tests/test_train_ops_gpu.py -x -q \
    -k "bf16_mode_tensor_core or varying_cloud or batchnorm" > gpurun_out/r2_sanitizer_racecheck.txt 2>&1
echo "racecheck rc=$?" >> gpurun_out/r2_sanitizer_racecheck.txt
du -sh gpurun_out; ls -la gpurun_out | tail -5
