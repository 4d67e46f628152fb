#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -q -k "voxelize or varying or ragged or smoke or plugin_through or bench_launch" > gpurun_out/r2_pytest_c.txt 2>&1; grep -E "passed|failed|^FAILED|^E  " gpurun_out/r2_pytest_c.txt | cut -c1-250
timeout 600 python bench.py --steps 50 --warmup 5 --no-extras --no-train --no-cpu-baseline > gpurun_out/r2_bench_assign.json 2> gpurun_out/r2_bench_assign.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_assign.json').read().strip().splitlines()[-1])
print(d['value'], [(h['kernel'][:20], h['us'], h['frac']) for h in d['roofline_hbm']])
PY
