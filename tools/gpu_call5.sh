#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_train_step.csv python tools/prof_train.py 4 > gpurun_out/r2_prof_train.log 2>&1
tail -3 gpurun_out/r2_prof_train.log; wc -l gpurun_out/r2_launches_train_step.csv
