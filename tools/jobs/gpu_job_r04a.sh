#!/bin/bash
# GPU job 4A: lagged form without the pass loop: 6 = no max ops (wrong), 8 = max on the current quarter, 9 = max on the next quarter after its load (8, 9: correct unless p overflows).
# experiment had), 3 = classic steps + those commits.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r04a_build.log 2>&1
: > gpurun_out/r04a_probe.jsonl
for rep in 1 2; do
for st in 0 6 8 9; do
X2I_ATTN_LAG=$st timeout 120 python tools/attn_probe.py --tag "lag=$st" >> gpurun_out/r04a_probe.jsonl 2>> gpurun_out/r04a_probe.err
done; done
cut -c1-330 gpurun_out/r04a_probe.jsonl; tail -3 gpurun_out/r04a_probe.err
