#!/bin/bash
# GPU job 4Q: (1) tests of the tree with 256-bit RoPE-table loads in the QKV epilogue and the shared 256-bit row store in the attention
# forward / backward epilogues; (2) gemm_probe of that tree and of two libraries that pace the epilogue's store stream (nanosleep 300 / 1000 ns
# between 32-column chunks; built from a tree with an X2I_EPI_SLEEP_NS macro in the generic epilogue, since removed).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_flux.py tests/test_gpu_bwd_kernels.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/r04q_tests.log 2>&1; echo "tests rc=$?" | tee gpurun_out/r04q_rc.log
tail -3 gpurun_out/r04q_tests.log
cp x2i_b200/libx2i_b200.so /tmp/keep.so
for v in keep sleep300 sleep1000; do
if [ $v = keep ]; then cp /tmp/keep.so x2i_b200/libx2i_b200.so; else cp libx2i_$v.so x2i_b200/libx2i_b200.so; fi
timeout 600 python tools/gemm_probe.py > gpurun_out/r04q_gemm_probe_$v.jsonl 2>> gpurun_out/r04q_gemm_probe.err
done
cp /tmp/keep.so x2i_b200/libx2i_b200.so
python - <<PY
import json
rows = [[json.loads(l) for l in open("gpurun_out/r04q_gemm_probe_%s.jsonl" % v)] for v in ("keep", "sleep300", "sleep1000")]
for x, y, z in zip(*rows):
    print(x["case"], "|", round(x["ms"], 4), round(x["tflops_sustained"], 1), "| 300 ns", round(y["ms"], 4), round(y["tflops_sustained"], 1), "| 1000 ns", round(z["ms"], 4), round(z["tflops_sustained"], 1))
PY
tail -3 gpurun_out/r04q_gemm_probe.err
