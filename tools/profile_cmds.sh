#!/bin/bash
# ncu evidence for the bench command (run under gpurun; outputs in gpurun_out/).  Never a bench number.
# X2I_NCU=1 makes bench.py bracket its timed region with cudaProfilerStart/Stop.
# Every .ncu-rep is exported to its raw-page CSV on the box and removed (gpurun_out/ is capped at 64 MiB); only the
# attention report itself travels back.
set -x
export X2I_NCU=1
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train --no-library-baseline"
full() {  # full <name> <kernel regex> <count> <extra ncu args...> -- <command...>
  local name=$1 regex=$2 count=$3; shift 3
  local extra=()
  while [ "$1" != "--" ]; do extra+=("$1"); shift; done
  shift
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$regex" -c "$count" "${extra[@]}" -o gpurun_out/$name -f "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  [ "$name" = prof_attn ] || rm -f gpurun_out/$name.ncu-rep
}
# 1) launch list: every kernel of the timed steps with its device time
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
if [ "$1" != "list-only" ]; then
# 2) full captures: fused attention kernel, the dominant GEMMs, the row-wise kernels
full prof_attn mmdit_attention 2 --profile-from-start off -- $BENCH
ncu -i gpurun_out/prof_attn.ncu-rep --page source --csv > gpurun_out/prof_attn_source.csv 2>/dev/null
if [ "$1" != "no-gemm" ]; then
full prof_gemm "gemm2?_tcgen05" 8 --profile-from-start off -- $BENCH
full prof_rowwise "ln_modulate|skinny" 4 --profile-from-start off -- $BENCH
fi
fi
if [ "$1" = "all" ]; then
# 3) VAE decode and the ControlNeXt nets: launch lists + full captures of the conv / GroupNorm kernels
unset X2I_NCU
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/vae_launches.csv python tools/bench_vae.py --steps 1 --warmup 1 > gpurun_out/ncu_vae.log 2>&1
full prof_vae "conv2d_tcgen05|gn_apply|gn_stats_partial|softmax_rows" 24 --launch-skip 70 -- python tools/bench_vae.py --steps 1 --warmup 0
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/cn_launches.csv python tools/profile_controlnext.py > gpurun_out/ncu_cn.log 2>&1
fi
du -sh gpurun_out; ls -la gpurun_out/
