#!/bin/bash
set -x
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q ) > gpurun_out/r2_pytest_full.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_full.txt
grep -E "passed|failed|^FAILED|^E  |real" gpurun_out/r2_pytest_full.txt | cut -c1-250
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2_smoke.txt 2>&1; tail -6 gpurun_out/r2_smoke.txt
( time timeout 900 python bench.py ) > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; tail -4 gpurun_out/r2_bench_default.err
( time timeout 600 python bench.py --impl reference ) > gpurun_out/r2_bench_ref_default.json 2> gpurun_out/r2_bench_ref_default.err; tail -4 gpurun_out/r2_bench_ref_default.err
