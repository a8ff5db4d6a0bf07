#!/bin/bash
# ncu evidence for the bench command (run under gpurun; outputs in gpurun_out/).  Never a bench number.
set -x
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
# 1) launch list: every kernel of two timed steps with its device time (skip the 3 warm-up steps)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 900 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
# 2) full capture of the fused attention kernel and of the dominant GEMM
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mmdit_attention -s 40 -c 2 -o gpurun_out/prof_attn -f $BENCH > gpurun_out/ncu_attn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 1500 -c 6 -o gpurun_out/prof_gemm -f $BENCH > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out/
