#!/bin/bash
# round-2 GPU job L: dQ launch of the attention backward with the owner tiles in TMEM (TS-form T products): parity, then sustained probe.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02l_build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_bwd_kernels.py -x -q -m gpu > gpurun_out/r02l_tests.log 2>&1; echo "tests rc=$?" | tee gpurun_out/r02l_rc.log
tail -3 gpurun_out/r02l_tests.log
: > gpurun_out/r02l_probe.jsonl
timeout 120 python tools/attn_probe.py --bwd --tag "bwd: dQ launch TS" >> gpurun_out/r02l_probe.jsonl 2>> gpurun_out/r02l_probe.err
timeout 120 python tools/attn_probe.py --bwd --B 2 --tag "bwd: dQ launch TS B=2" >> gpurun_out/r02l_probe.jsonl 2>> gpurun_out/r02l_probe.err
timeout 200 python tools/gpu_check.py --one perf_bwd 2>&1 | grep RESULT > gpurun_out/r02l_perf_bwd.txt
cat gpurun_out/r02l_probe.jsonl gpurun_out/r02l_perf_bwd.txt; tail -3 gpurun_out/r02l_probe.err
