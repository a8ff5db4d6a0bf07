#!/bin/bash
# GPU job 4M: what part of the epilogue is exposed?  Libraries whose CTA-pair kernel epilogue only READS the accumulator tile from TMEM
# (-DX2I_GEMM_EPI_TMEM_ONLY) or only WRITES the output (-DX2I_GEMM_EPI_STORE_ONLY), against the shipped one (no results in either).
mkdir -p gpurun_out
cp x2i_b200/libx2i_b200.so /tmp/keep.so
for v in TMEM_ONLY STORE_ONLY; do
cp libx2i_$v.so x2i_b200/libx2i_b200.so
timeout 600 python tools/gemm_probe.py > gpurun_out/r04m_gemm_probe_$v.jsonl 2>> gpurun_out/r04m_gemm_probe.err
done
cp /tmp/keep.so x2i_b200/libx2i_b200.so
python - <<PY
import json
a = [json.loads(l) for l in open("gpurun_out/r04m_gemm_probe_TMEM_ONLY.jsonl")]
b = [json.loads(l) for l in open("gpurun_out/r04m_gemm_probe_STORE_ONLY.jsonl")]
for x, y in zip(a, b):
    print(x["case"], "| tmem read only", round(x["ms"], 4), "ms", round(x["tflops_sustained"], 1), "TF", x["sm_mhz_median"], "| store only", round(y["ms"], 4), "ms", round(y["tflops_sustained"], 1), "TF", y["sm_mhz_median"])
PY
tail -3 gpurun_out/r04m_gemm_probe.err
