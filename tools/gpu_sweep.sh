#!/bin/bash
# gpurun helper: scenes-per-launch sweep of the inference bench (SURVEY 8d timing protocol), one JSON line per batch size
set -x
mkdir -p gpurun_out
: > gpurun_out/r2_bench_sweep.jsonl
for b in 1 2 4 6 8 12; do
  timeout 300 python bench.py --scenes-per-step $b --steps 40 --warmup 5 --no-extras --no-train --no-cpu-baseline >> gpurun_out/r2_bench_sweep.jsonl 2>> gpurun_out/r2_bench_sweep.err
done
python - <<'PY'
import json
for l in open('gpurun_out/r2_bench_sweep.jsonl'):
    d = json.loads(l)
    print(d['config'].get('scenes_per_step'), round(d['value'], 1), round(d['ms_per_step'], 3), round(d['e2e']['value'], 1), round(d['roofline']['frac'], 3), d['clocks']['sm_mhz'])
PY
