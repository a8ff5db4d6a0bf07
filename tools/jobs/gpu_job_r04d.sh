#!/bin/bash
# GPU job 4D: lagged form with the redo pass behind a noinline call: correctness, then sustained A/B.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r04d_build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "lagged_form or forms_are_bit_identical" > gpurun_out/r04d_tests.log 2>&1; echo "tests rc=$?" | tee gpurun_out/r04d_rc.log
tail -12 gpurun_out/r04d_tests.log
: > gpurun_out/r04d_probe.jsonl
for rep in 1 2; do
X2I_ATTN_LAG=0 timeout 120 python tools/attn_probe.py --tag "persistent (default)" >> gpurun_out/r04d_probe.jsonl 2>> gpurun_out/r04d_probe.err
X2I_ATTN_LAG=1 timeout 120 python tools/attn_probe.py --tag "lagged-max" >> gpurun_out/r04d_probe.jsonl 2>> gpurun_out/r04d_probe.err
done
timeout 120 python tools/attn_probe.py --sdpa --tag "sdpa" >> gpurun_out/r04d_probe.jsonl 2>> gpurun_out/r04d_probe.err
cut -c1-360 gpurun_out/r04d_probe.jsonl; tail -3 gpurun_out/r04d_probe.err
