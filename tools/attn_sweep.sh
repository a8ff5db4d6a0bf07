#!/bin/bash
# Times the attention kernel variants (env-selected instantiations) in separate processes; output in gpurun_out/attn_sweep.log
# Needs a library with the experiment instantiations: X2I_BUILD_EXPERIMENTS=1 python __graft_entry__.py --force  (rebuild without it afterwards)
out=gpurun_out/attn_sweep.log; : > $out
run() { echo "### $*" >> $out; env "$@" timeout 120 python tools/gpu_check.py --one attn_variants 2>&1 | grep -E "RESULT|ATTTRACE|rror" >> $out; }
if [ $# -eq 0 ]; then set -- "X2I_ATTN_POLY8=0"; fi
for v in "$@"; do run $v; done
cat $out
