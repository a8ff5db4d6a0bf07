#!/bin/bash
# GPU job 4C: cost of the polynomial-lane guard in the lagged form: 1 = one chain, 2 = no guard (unsafe), 3 = two chains.
# experiment had), 3 = classic steps + those commits.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r04c_build.log 2>&1
: > gpurun_out/r04c_probe.jsonl
for rep in 1 2; do
for st in 0 1 2 3; do
X2I_ATTN_LAG=$st timeout 120 python tools/attn_probe.py --tag "lag=$st" >> gpurun_out/r04c_probe.jsonl 2>> gpurun_out/r04c_probe.err
done; done
cut -c1-330 gpurun_out/r04c_probe.jsonl; tail -3 gpurun_out/r04c_probe.err
