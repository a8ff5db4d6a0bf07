#!/bin/bash
# round-2 GPU job D: MLLM prefill (N3) kernels + drop-in, the attention kernel after its loop generalisation, pair/non-pair bit identity.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02d_build.log 2>&1
timeout 1200 python -m pytest tests/test_mllm_prefill.py -x -q -m gpu -s > gpurun_out/r02d_mllm.log 2>&1; echo "mllm rc=$?" | tee gpurun_out/r02d_rc.log
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py tests/test_gpu_bwd_kernels.py tests/test_gpu_flux.py -q -m gpu > gpurun_out/r02d_attn_regress.log 2>&1; echo "regress rc=$?" | tee -a gpurun_out/r02d_rc.log
X2I_ATTN_PAIR=0 timeout 120 python tools/attn_probe.py --tag "single after loop generalisation" > gpurun_out/r02d_probe.jsonl 2>> gpurun_out/r02d_probe.err
grep -v "^$" gpurun_out/r02d_mllm.log | tail -40; tail -5 gpurun_out/r02d_attn_regress.log; cat gpurun_out/r02d_probe.jsonl
