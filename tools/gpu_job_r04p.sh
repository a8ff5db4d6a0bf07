#!/bin/bash
# GPU job 4P: does pacing the epilogue's store stream (nanosleep between 32-column chunks) reduce its interference with the mainloop?
mkdir -p gpurun_out
cp x2i_b200/libx2i_b200.so /tmp/keep.so
for v in keep sleep300 sleep1000; do
if [ $v = keep ]; then cp /tmp/keep.so x2i_b200/libx2i_b200.so; else cp libx2i_$v.so x2i_b200/libx2i_b200.so; fi
timeout 600 python tools/gemm_probe.py > gpurun_out/r04p_gemm_probe_$v.jsonl 2>> gpurun_out/r04p_gemm_probe.err
done
cp /tmp/keep.so x2i_b200/libx2i_b200.so
python - <<PY
import json
rows = [[json.loads(l) for l in open("gpurun_out/r04p_gemm_probe_%s.jsonl" % v)] for v in ("keep", "sleep300", "sleep1000")]
for x, y, z in zip(*rows):
    print(x["case"], "|", round(x["ms"], 4), round(x["tflops_sustained"], 1), "| 300 ns", round(y["ms"], 4), round(y["tflops_sustained"], 1), "| 1000 ns", round(z["ms"], 4), round(z["tflops_sustained"], 1))
PY
tail -3 gpurun_out/r04p_gemm_probe.err
