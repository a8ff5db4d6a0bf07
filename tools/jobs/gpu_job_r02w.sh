#!/bin/bash
# round-2 GPU job W: tensor-core layer-mixing convolution: which tile geometry / descriptor form the hardware accepts, parity, timing.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02w_build.log 2>&1
: > gpurun_out/r02w_probe.jsonl
for al in 0 1; do for bo in 1 0; do
X2I_PROJCONV_ALIGNED=$al X2I_PROJCONV_BASE_OFFSET=$bo timeout 120 python tools/probe_projconv.py >> gpurun_out/r02w_probe.jsonl 2>> gpurun_out/r02w_probe.err; echo "aligned=$al base_offset=$bo rc=$?" | tee -a gpurun_out/r02w_rc.log
done; done
cat gpurun_out/r02w_probe.jsonl; tail -5 gpurun_out/r02w_probe.err
