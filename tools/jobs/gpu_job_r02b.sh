#!/bin/bash
# round-2 GPU job B: the CTA-pair attention forward -- parity tests, then the sustained-load probe of pair / single / SDPA.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02b_build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py tests/test_gpu_bwd_kernels.py -x -q -m gpu -k "attention" > gpurun_out/r02b_attn_tests.log 2>&1; echo "attn tests rc=$?" | tee gpurun_out/r02b_rc.log
tail -3 gpurun_out/r02b_attn_tests.log
: > gpurun_out/r02b_attn_probe.jsonl
X2I_ATTN_PAIR=1 timeout 120 python tools/attn_probe.py >> gpurun_out/r02b_attn_probe.jsonl 2>> gpurun_out/r02b_probe.err
X2I_ATTN_PAIR=0 timeout 120 python tools/attn_probe.py >> gpurun_out/r02b_attn_probe.jsonl 2>> gpurun_out/r02b_probe.err
timeout 120 python tools/attn_probe.py --sdpa >> gpurun_out/r02b_attn_probe.jsonl 2>> gpurun_out/r02b_probe.err
X2I_ATTN_PAIR=1 timeout 120 python tools/attn_probe.py --B 2 >> gpurun_out/r02b_attn_probe.jsonl 2>> gpurun_out/r02b_probe.err
timeout 120 python tools/attn_probe.py --bwd >> gpurun_out/r02b_attn_probe.jsonl 2>> gpurun_out/r02b_probe.err
timeout 120 python tools/attn_probe.py --bwd --sdpa >> gpurun_out/r02b_attn_probe.jsonl 2>> gpurun_out/r02b_probe.err
cat gpurun_out/r02b_attn_probe.jsonl; tail -5 gpurun_out/r02b_probe.err
