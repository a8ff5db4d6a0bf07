#!/bin/bash
# GPU job 4N: 256-bit stores in the GEMM epilogues: correctness (kernel + flux tests), then gemm_probe.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r04n_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_flux.py -x -q -m gpu > gpurun_out/r04n_tests.log 2>&1; echo "tests rc=$?" | tee gpurun_out/r04n_rc.log
tail -3 gpurun_out/r04n_tests.log
timeout 600 python tools/gemm_probe.py > gpurun_out/r04n_gemm_probe.jsonl 2> gpurun_out/r04n_gemm_probe.err
python - <<PY
import json
for l in open("gpurun_out/r04n_gemm_probe.jsonl"):
    j = json.loads(l)
    print(j["case"], "|", round(j["ms"], 4), "ms", round(j["tflops_sustained"], 1), "TF", j["sm_mhz_median"], "MHz", round(j["tensor_util_at_clock"], 3))
PY
tail -3 gpurun_out/r04n_gemm_probe.err
