#!/bin/bash
# ncu launch list of one distillation train step on a reduced block count (same kernels, same shapes per block).
set -x
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv \
  python tools/bench_train.py --batch 1 --steps 1 --warmup 1 --layers 2 4 > gpurun_out/ncu_train_launches.log 2>&1
tail -2 gpurun_out/ncu_train_launches.log
