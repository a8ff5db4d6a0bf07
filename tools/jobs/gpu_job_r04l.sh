#!/bin/bash
# GPU job 4L: how much of each GEMM is exposed epilogue?  gemm_probe with the shipped library and with one whose CTA-pair kernel skips
# the epilogue (-DX2I_GEMM_SKIP_EPI, no results).
mkdir -p gpurun_out
timeout 600 python tools/gemm_probe.py > gpurun_out/r04l_gemm_probe.jsonl 2> gpurun_out/r04l_gemm_probe.err
cp x2i_b200/libx2i_b200.so /tmp/keep.so; cp libx2i_noepi.so x2i_b200/libx2i_b200.so
timeout 600 python tools/gemm_probe.py > gpurun_out/r04l_gemm_probe_noepi.jsonl 2>> gpurun_out/r04l_gemm_probe.err
cp /tmp/keep.so x2i_b200/libx2i_b200.so
python - <<PY
import json
a = [json.loads(l) for l in open("gpurun_out/r04l_gemm_probe.jsonl")]
b = [json.loads(l) for l in open("gpurun_out/r04l_gemm_probe_noepi.jsonl")]
for x, y in zip(a, b):
    print(x["case"], "|", round(x["ms"], 4), "ms", round(x["tflops_sustained"], 1), "TF", x["sm_mhz_median"], "| no epilogue", round(y["ms"], 4), "ms", round(y["tflops_sustained"], 1), "TF", y["sm_mhz_median"], "| exposed us", round((x["ms"] - y["ms"]) * 1e3, 1))
PY
tail -3 gpurun_out/r04l_gemm_probe.err
