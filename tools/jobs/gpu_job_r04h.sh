#!/bin/bash
# GPU job 4H: column-split + lagged steps + cross-tile prefetch of the next tile's scores (X2I_ATTN_CS=2): predicated loads (shipped .so) and the
# branch form (alt .so built with -DX2I_CS_PREFETCH_BRANCH), against the default.
mkdir -p gpurun_out
: > gpurun_out/r04h_probe.jsonl
for rep in 1 2; do
for cs in 0 2; do
X2I_ATTN_CS=$cs timeout 120 python tools/attn_probe.py --tag "cs=$cs pred" >> gpurun_out/r04h_probe.jsonl 2>> gpurun_out/r04h_probe.err
done; done
cp x2i_b200/libx2i_b200.so /tmp/keep.so; cp libx2i_alt.so x2i_b200/libx2i_b200.so
for rep in 1 2; do
X2I_ATTN_CS=2 timeout 120 python tools/attn_probe.py --tag "cs=2 branch" >> gpurun_out/r04h_probe.jsonl 2>> gpurun_out/r04h_probe.err
done
cp /tmp/keep.so x2i_b200/libx2i_b200.so
python - <<PY
import json
for l in open("gpurun_out/r04h_probe.jsonl"):
    j = json.loads(l)
    print(j["variant"], round(j["tflops_sustained"], 1), round(j["tflops_first20"], 1), j["sm_mhz_median"], round(j["tensor_util_at_clock"], 3), j.get("rel_err_vs_fp32_sdpa"))
PY
tail -3 gpurun_out/r04h_probe.err
