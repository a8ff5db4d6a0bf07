#!/bin/bash
# Timing experiments on the attention backward (needs X2I_BUILD_EXPERIMENTS=1 python __graft_entry__.py --force)
out=gpurun_out/attn_bwd_sweep.log; : > $out
run() { echo "### $*" >> $out; env "$@" timeout 200 python tools/gpu_check.py --one perf_bwd 2>&1 | grep RESULT | python -c "import sys,json; j=json.loads(sys.stdin.read()[7:]); print({k:round(v['ms'],4) for k,v in j.items() if k.startswith('attn')})" >> $out; }
run X2I_ATTN_BWD_ONLY=0
run X2I_ATTN_BWD_ONLY=1
run X2I_ATTN_BWD_ONLY=2
run X2I_ATTN_BWD_ONLY=1 X2I_ATTN_BWD_DBG=1
run X2I_ATTN_BWD_ONLY=2 X2I_ATTN_BWD_DBG=1
cat $out
