#!/bin/bash
# round-2 GPU job M: attention backward with column-split soft-max warpgroups + separate T1 / T2 barriers: parity, sustained probe, per-launch times.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02m_build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_bwd_kernels.py -x -q -m gpu > gpurun_out/r02m_tests.log 2>&1; echo "tests rc=$?" | tee gpurun_out/r02m_rc.log
tail -3 gpurun_out/r02m_tests.log
: > gpurun_out/r02m_probe.jsonl
timeout 120 python tools/attn_probe.py --bwd --tag "bwd: split warpgroups" >> gpurun_out/r02m_probe.jsonl 2>> gpurun_out/r02m_probe.err
timeout 120 python tools/attn_probe.py --bwd --sdpa --tag "sdpa fwd+bwd" >> gpurun_out/r02m_probe.jsonl 2>> gpurun_out/r02m_probe.err
timeout 120 python tools/attn_probe.py --sdpa --tag "sdpa fwd" >> gpurun_out/r02m_probe.jsonl 2>> gpurun_out/r02m_probe.err
cat gpurun_out/r02m_probe.jsonl; tail -3 gpurun_out/r02m_probe.err
X2I_BUILD_EXPERIMENTS=1 python __graft_entry__.py --force > gpurun_out/r02m_build_exp.log 2>&1
bash tools/attn_bwd_sweep.sh > /dev/null 2>&1; cp gpurun_out/attn_bwd_sweep.log gpurun_out/r02m_bwd_sweep.log; cat gpurun_out/r02m_bwd_sweep.log
