#!/bin/bash
# ncu evidence for the distillation train step (run under gpurun; outputs in gpurun_out/).  Never a bench number.
set -x
# 1) launch list of one train step on a reduced block count (same kernels and shapes per block)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv \
  python tools/bench_train.py --batch 1 --steps 1 --warmup 1 --layers 2 4 > gpurun_out/ncu_train_launches.log 2>&1
# 2) full captures of every backward kernel at its FLUX shape
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/prof_bwd -f \
  python tools/profile_bwd_kernels.py > gpurun_out/ncu_bwd.log 2>&1
tail -2 gpurun_out/ncu_bwd.log
