#!/bin/bash
# GPU job 3Z: bisect of the lagged form (all without max ops, wrong results): 4 = with redo pass loop, 6 = no pass loop, 7 = no pass loop and no top-of-step rescale block.
# experiment had), 3 = classic steps + those commits.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r03z_build.log 2>&1
: > gpurun_out/r03z_probe.jsonl
for rep in 1 2; do
for st in 0 4 6 7; do
X2I_ATTN_LAG=$st timeout 120 python tools/attn_probe.py --tag "lag=$st" >> gpurun_out/r03z_probe.jsonl 2>> gpurun_out/r03z_probe.err
done; done
cut -c1-330 gpurun_out/r03z_probe.jsonl; tail -3 gpurun_out/r03z_probe.err
