#!/bin/bash
# GPU job 4K: the CTA-pair GEMM by epilogue kind at the denoise step's shapes, sustained load (is the QKV epilogue visible?).
mkdir -p gpurun_out
timeout 600 python tools/gemm_probe.py > gpurun_out/r04k_gemm_probe.jsonl 2> gpurun_out/r04k_gemm_probe.err
python - <<PY
import json
for l in open("gpurun_out/r04k_gemm_probe.jsonl"):
    j = json.loads(l)
    print(j["case"], "|", round(j["ms"], 4), "ms", round(j["tflops_sustained"], 1), "TF", j["sm_mhz_median"], "MHz", round(j["tensor_util_at_clock"], 3))
PY
tail -3 gpurun_out/r04k_gemm_probe.err
