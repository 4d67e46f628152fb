#!/bin/bash
set -x
mkdir -p gpurun_out
run() { # name, kernel regex, skip, count
  timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:$2 --launch-skip $3 -c $4 -f -o /tmp/$1 python tools/prof_train.py 4 > gpurun_out/r2_ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/r2_ncu_$1_raw.csv 2>/dev/null
}
run tr_bn_apply bn_apply_kernel 0 3
run tr_bn_bwd_apply bn_bwd_apply_kernel 34 4
run tr_wgrad_l0 wgrad_tc_kernel 34 4
run tr_wgrad_shrink wgrad_tc_kernel 1 2
ls -la gpurun_out | tail -8
