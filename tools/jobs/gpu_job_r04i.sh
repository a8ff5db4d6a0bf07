#!/bin/bash
# GPU job 4I: share of the soft-max exponentials on the FMA pipe (POLY8 of every 8 pairs) re-tuned for the lagged form: libraries built with
# -DX2I_ATT_POLY8=1 / 3 against the shipped 2.
mkdir -p gpurun_out
: > gpurun_out/r04i_probe.jsonl
cp x2i_b200/libx2i_b200.so /tmp/keep.so
for rep in 1 2; do
for P in 2 1 3; do
if [ $P = 2 ]; then cp /tmp/keep.so x2i_b200/libx2i_b200.so; else cp libx2i_poly$P.so x2i_b200/libx2i_b200.so; fi
timeout 120 python tools/attn_probe.py --tag "lagged POLY8=$P" >> gpurun_out/r04i_probe.jsonl 2>> gpurun_out/r04i_probe.err
done; done
cp /tmp/keep.so x2i_b200/libx2i_b200.so
python - <<PY
import json
for l in open("gpurun_out/r04i_probe.jsonl"):
    j = json.loads(l)
    print(j["variant"], round(j["tflops_sustained"], 1), round(j["tflops_first20"], 1), j["sm_mhz_median"], round(j["tensor_util_at_clock"], 3), j.get("rel_err_vs_fp32_sdpa"))
PY
tail -3 gpurun_out/r04i_probe.err
