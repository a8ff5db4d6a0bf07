#!/bin/bash
# GPU job 3S: modulation GEMV of blocks 1.. on a side stream next to block 0: parity (graph == eager, full model), A/B of the bench line.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r03s_build.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_flux.py tests/test_gpu_parity_full.py tests/test_gpu_train.py tests/test_controlnext.py -x -q -m gpu > gpurun_out/r03s_tests.log 2>&1; echo "tests rc=$?" | tee gpurun_out/r03s_rc.log
tail -3 gpurun_out/r03s_tests.log
for rep in 1 2; do for ov in 0 1; do
X2I_OVERLAP_MOD=$ov timeout 600 python bench.py --steps 20 --warmup 5 --no-train --no-cpu-baseline --no-library-baseline > gpurun_out/r03s_b.json 2> gpurun_out/r03s_b.err; echo "overlap=$ov rep=$rep rc=$?"
python - <<PY
import json
j = json.loads([l for l in open("gpurun_out/r03s_b.json") if l.startswith("{")][0])
r = j["roofline"]
print("   value", round(j["value"], 3), "ms", round(j["ms_per_step"], 3), "e2e", round(j["e2e"]["value"], 3), "attn ms", round(r["ms_per_launch"], 4), j["clocks"]["sm_mhz"])
PY
done; done
