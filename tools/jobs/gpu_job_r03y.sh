#!/bin/bash
# GPU job 3Y: lagged-max form: 4 = no max ops at all (wrong, upper bound), 5 = max on the next quarter right after its load.
# experiment had), 3 = classic steps + those commits.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r03y_build.log 2>&1
: > gpurun_out/r03y_probe.jsonl
for rep in 1 2; do
for st in 0 1 4 5; do
X2I_ATTN_LAG=$st timeout 120 python tools/attn_probe.py --tag "lag=$st" >> gpurun_out/r03y_probe.jsonl 2>> gpurun_out/r03y_probe.err
done; done
cut -c1-330 gpurun_out/r03y_probe.jsonl; tail -3 gpurun_out/r03y_probe.err
