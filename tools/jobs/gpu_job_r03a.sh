#!/bin/bash
# GPU job 3A: dK/dV launch with 32-query stream tiles, all products TS: parity, then A/B on one box.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r03a_build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_bwd_kernels.py -x -q -m gpu > gpurun_out/r03a_tests.log 2>&1; echo "tests rc=$?" | tee gpurun_out/r03a_rc.log
tail -12 gpurun_out/r03a_tests.log
: > gpurun_out/r03a_probe.jsonl
for rep in 1 2; do
X2I_ATTN_BWD_KV32=0 timeout 120 python tools/attn_probe.py --bwd --tag "bwd: dK/dV 64-query SS form" >> gpurun_out/r03a_probe.jsonl 2>> gpurun_out/r03a_probe.err
X2I_ATTN_BWD_KV32=1 timeout 120 python tools/attn_probe.py --bwd --tag "bwd: dK/dV 32-query TS form" >> gpurun_out/r03a_probe.jsonl 2>> gpurun_out/r03a_probe.err
done
timeout 120 python tools/attn_probe.py --bwd --sdpa --tag "sdpa fwd+bwd" >> gpurun_out/r03a_probe.jsonl 2>> gpurun_out/r03a_probe.err
timeout 120 python tools/attn_probe.py --sdpa --tag "sdpa fwd" >> gpurun_out/r03a_probe.jsonl 2>> gpurun_out/r03a_probe.err
X2I_ATTN_BWD_KV32=1 timeout 120 python tools/attn_probe.py --bwd --B 2 --tag "bwd: 32-query TS form B=2" >> gpurun_out/r03a_probe.jsonl 2>> gpurun_out/r03a_probe.err
cut -c1-330 gpurun_out/r03a_probe.jsonl; tail -3 gpurun_out/r03a_probe.err
