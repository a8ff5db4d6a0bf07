#!/bin/bash
# GPU job 4G: column-split soft-max with lagged steps (X2I_ATTN_CS=2; redo pass not wired yet): sustained A/B against the default (lagged,
# one warpgroup per tile) and the classic column-split form.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r04g_build.log 2>&1
: > gpurun_out/r04g_probe.jsonl
for rep in 1 2; do
for cs in 0 1 2; do
X2I_ATTN_CS=$cs timeout 120 python tools/attn_probe.py --tag "cs=$cs" >> gpurun_out/r04g_probe.jsonl 2>> gpurun_out/r04g_probe.err
done; done
timeout 120 python tools/attn_probe.py --sdpa --tag "sdpa" >> gpurun_out/r04g_probe.jsonl 2>> gpurun_out/r04g_probe.err
python - <<PY
import json
for l in open("gpurun_out/r04g_probe.jsonl"):
    j = json.loads(l)
    print(j["variant"], round(j["tflops_sustained"], 1), round(j["tflops_first20"], 1), j["sm_mhz_median"], round(j["tensor_util_at_clock"], 3), j.get("rel_err_vs_fp32_sdpa"))
PY
tail -3 gpurun_out/r04g_probe.err
