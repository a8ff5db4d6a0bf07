#!/bin/bash
# GPU job 3V: timing experiment -- stale-max steps WITHOUT the slow-path branch (X2I_ATTN_STALE=2, wrong when a row's max grows):
# is the 4 % loss of the stale form the basic-block boundary at every quarter?
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r03v_build.log 2>&1
: > gpurun_out/r03v_probe.jsonl
for rep in 1 2; do
for st in 0 1 2; do
X2I_ATTN_STALE=$st timeout 120 python tools/attn_probe.py --tag "stale=$st" >> gpurun_out/r03v_probe.jsonl 2>> gpurun_out/r03v_probe.err
done; done
cut -c1-330 gpurun_out/r03v_probe.jsonl; tail -3 gpurun_out/r03v_probe.err
