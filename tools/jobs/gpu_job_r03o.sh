#!/bin/bash
# GPU job 3O: column-split soft-max form of the persistent attention forward: correctness, then sustained A/B against the default and SDPA.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r03o_build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "forms_are_bit_identical" > gpurun_out/r03o_tests.log 2>&1; echo "tests rc=$?" | tee gpurun_out/r03o_rc.log
tail -8 gpurun_out/r03o_tests.log
: > gpurun_out/r03o_probe.jsonl
for rep in 1 2; do
X2I_ATTN_CS=0 timeout 120 python tools/attn_probe.py --tag "persistent (default)" >> gpurun_out/r03o_probe.jsonl 2>> gpurun_out/r03o_probe.err
X2I_ATTN_CS=1 timeout 120 python tools/attn_probe.py --tag "column-split soft-max" >> gpurun_out/r03o_probe.jsonl 2>> gpurun_out/r03o_probe.err
done
timeout 120 python tools/attn_probe.py --sdpa --tag "sdpa" >> gpurun_out/r03o_probe.jsonl 2>> gpurun_out/r03o_probe.err
cut -c1-360 gpurun_out/r03o_probe.jsonl; tail -3 gpurun_out/r03o_probe.err
