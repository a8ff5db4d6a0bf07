#!/bin/bash
# ncu evidence for the bench command (run under gpurun; outputs in gpurun_out/).  Never a bench number.
# X2I_NCU=1 makes bench.py bracket its timed region with cudaProfilerStart/Stop.
set -x
export X2I_NCU=1
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train"
# 1) launch list: every kernel of the timed steps with its device time
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
# 2) full captures: fused attention kernel, the dominant GEMMs, the row-wise kernels
if [ "$1" != "list-only" ]; then
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:mmdit_attention -c 2 -o gpurun_out/prof_attn -f $BENCH > gpurun_out/ncu_attn.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"gemm2?_tcgen05" -c 10 -o gpurun_out/prof_gemm -f $BENCH > gpurun_out/ncu_gemm.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"ln_modulate|skinny" -c 4 -o gpurun_out/prof_rowwise -f $BENCH > gpurun_out/ncu_rowwise.log 2>&1
fi
ls -la gpurun_out/
