#!/bin/bash
# GPU job 3X: why is the lagged-max form slow?  LAG=1 as designed, 2 = + one tcgen05.commit behind every P.V quarter (as the stale-max
# experiment had), 3 = classic steps + those commits.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r03x_build.log 2>&1
: > gpurun_out/r03x_probe.jsonl
for rep in 1 2; do
for st in 0 1 2 3; do
X2I_ATTN_LAG=$st timeout 120 python tools/attn_probe.py --tag "lag=$st" >> gpurun_out/r03x_probe.jsonl 2>> gpurun_out/r03x_probe.err
done; done
cut -c1-330 gpurun_out/r03x_probe.jsonl; tail -3 gpurun_out/r03x_probe.err
