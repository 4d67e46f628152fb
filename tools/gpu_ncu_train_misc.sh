#!/bin/bash
# gpurun helper: ncu --set full of the non-GEMM, non-BatchNorm kernels of one training iteration
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --profile-from-start off --clock-control none -k regex:'pfn_|fuse_bwd|grad_combine|pack_weights|heads_grad|adam|loss_|canvas|wgrad' -c 60 -f -o /tmp/train_misc python tools/prof_train.py 4 > gpurun_out/ncu_train_misc.log 2>&1
ncu -i /tmp/train_misc.ncu-rep --page raw --csv > gpurun_out/ncu_train_misc_raw.csv 2>/dev/null
ls -la gpurun_out/ncu_train_misc_raw.csv
