#!/bin/bash
# round-2 GPU job O: attention backward as 2-CTA clusters sharing the streamed tiles (TMA multicast): bit identity, then A/B on one box.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02o_build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_bwd_kernels.py -x -q -m gpu > gpurun_out/r02o_tests.log 2>&1; echo "tests rc=$?" | tee gpurun_out/r02o_rc.log
tail -5 gpurun_out/r02o_tests.log
: > gpurun_out/r02o_probe.jsonl
for rep in 1 2; do
X2I_ATTN_BWD_MC=0 timeout 120 python tools/attn_probe.py --bwd --tag "bwd: one CTA per tile" >> gpurun_out/r02o_probe.jsonl 2>> gpurun_out/r02o_probe.err
X2I_ATTN_BWD_MC=1 timeout 120 python tools/attn_probe.py --bwd --tag "bwd: cluster multicast" >> gpurun_out/r02o_probe.jsonl 2>> gpurun_out/r02o_probe.err
done
timeout 120 python tools/attn_probe.py --bwd --sdpa --tag "sdpa fwd+bwd" >> gpurun_out/r02o_probe.jsonl 2>> gpurun_out/r02o_probe.err
timeout 120 python tools/attn_probe.py --sdpa --tag "sdpa fwd" >> gpurun_out/r02o_probe.jsonl 2>> gpurun_out/r02o_probe.err
X2I_ATTN_BWD_MC=1 timeout 120 python tools/attn_probe.py --bwd --B 2 --tag "bwd: cluster multicast B=2" >> gpurun_out/r02o_probe.jsonl 2>> gpurun_out/r02o_probe.err
cut -c1-330 gpurun_out/r02o_probe.jsonl; tail -3 gpurun_out/r02o_probe.err
