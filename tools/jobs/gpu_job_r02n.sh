#!/bin/bash
# round-2 GPU job N: attention backward, 6-stage ring in the dQ launch: parity + per-launch timing floor (experiments build).
mkdir -p gpurun_out
X2I_BUILD_EXPERIMENTS=1 python __graft_entry__.py --force > gpurun_out/r02n_build_exp.log 2>&1
timeout 600 python -m pytest tests/test_gpu_bwd_kernels.py -x -q -m gpu > gpurun_out/r02n_tests.log 2>&1; echo "tests rc=$?" | tee gpurun_out/r02n_rc.log
tail -3 gpurun_out/r02n_tests.log
bash tools/attn_bwd_sweep.sh > /dev/null 2>&1; cp gpurun_out/attn_bwd_sweep.log gpurun_out/r02n_bwd_sweep.log; cat gpurun_out/r02n_bwd_sweep.log
timeout 120 python tools/attn_probe.py --bwd --tag "bwd: 6-stage dQ ring" > gpurun_out/r02n_probe.jsonl 2> gpurun_out/r02n_probe.err; cat gpurun_out/r02n_probe.jsonl
