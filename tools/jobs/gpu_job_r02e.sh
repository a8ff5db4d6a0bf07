#!/bin/bash
# round-2 GPU job E: persistent attention kernel -- bit identity, lse path, then sustained probe (one-item vs persistent vs SDPA).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02e_build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_mllm_prefill.py tests/test_gpu_kernels.py -x -q -m gpu -k "attention or prefill" > gpurun_out/r02e_tests.log 2>&1; echo "tests rc=$?" | tee gpurun_out/r02e_rc.log
X2I_ATTN_PERSIST=1 timeout 600 python -m pytest tests/test_gpu_bwd_kernels.py tests/test_gpu_flux.py tests/test_gpu_train.py -x -q -m gpu > gpurun_out/r02e_tests_persist.log 2>&1; echo "persist-mode tests rc=$?" | tee -a gpurun_out/r02e_rc.log
tail -3 gpurun_out/r02e_tests.log; tail -3 gpurun_out/r02e_tests_persist.log
: > gpurun_out/r02e_probe.jsonl
X2I_ATTN_PERSIST=0 timeout 120 python tools/attn_probe.py --tag "one-item kernel" >> gpurun_out/r02e_probe.jsonl 2>> gpurun_out/r02e_probe.err
X2I_ATTN_PERSIST=1 timeout 120 python tools/attn_probe.py --tag "persistent kernel" >> gpurun_out/r02e_probe.jsonl 2>> gpurun_out/r02e_probe.err
timeout 120 python tools/attn_probe.py --sdpa >> gpurun_out/r02e_probe.jsonl 2>> gpurun_out/r02e_probe.err
X2I_ATTN_PERSIST=0 timeout 120 python tools/attn_probe.py --B 2 --tag "one-item kernel B=2" >> gpurun_out/r02e_probe.jsonl 2>> gpurun_out/r02e_probe.err
X2I_ATTN_PERSIST=1 timeout 120 python tools/attn_probe.py --B 2 --tag "persistent kernel B=2" >> gpurun_out/r02e_probe.jsonl 2>> gpurun_out/r02e_probe.err
cat gpurun_out/r02e_probe.jsonl; tail -3 gpurun_out/r02e_probe.err
