#!/bin/bash
# GPU job 3G: PDL experiment: does the event-bracketed in-step attention time change because the attention kernel itself is a dependent launch?
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r03g_build.log 2>&1
for cfg in "0 1" "1 1" "1 0" "0 1" "1 1" "1 0"; do set -- $cfg
X2I_PDL=$1 X2I_PDL_ATTN=$2 timeout 600 python bench.py --steps 20 --warmup 5 --no-train --no-cpu-baseline --no-library-baseline > gpurun_out/r03g_b.json 2> gpurun_out/r03g_b.err; echo "pdl=$1 pdl_attn=$2 rc=$?"
python - <<PY
import json
j = json.loads([l for l in open("gpurun_out/r03g_b.json") if l.startswith("{")][0])
r = j["roofline"]
print("   value", round(j["value"], 3), "ms", round(j["ms_per_step"], 3), "e2e", round(j["e2e"]["value"], 3), "attn ms", round(r["ms_per_launch"], 4), "iso", round(r["isolated_tflops"]), j["clocks"]["sm_mhz"])
PY
done
