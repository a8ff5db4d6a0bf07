#!/bin/bash
# GPU job 4V: opt-in experiment build -DX2I_EPI_STAGE (generic GEMM epilogue stores through a shared-memory transpose: 8 rows x one full
# 128-byte line per 256-bit store instruction): kernel tests with that library, then gemm_probe against the shipped one.
mkdir -p gpurun_out
cp x2i_b200/libx2i_b200.so /tmp/keep.so; cp libx2i_stage.so x2i_b200/libx2i_b200.so
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > gpurun_out/r04v_tests.log 2>&1; echo "tests (stage lib) rc=$?" | tee gpurun_out/r04v_rc.log
tail -2 gpurun_out/r04v_tests.log
timeout 300 python tools/gemm_probe.py --seconds 1.0 > gpurun_out/r04v_gemm_probe_stage.jsonl 2>> gpurun_out/r04v_gemm_probe.err
cp /tmp/keep.so x2i_b200/libx2i_b200.so
timeout 300 python tools/gemm_probe.py --seconds 1.0 > gpurun_out/r04v_gemm_probe_keep.jsonl 2>> gpurun_out/r04v_gemm_probe.err
python - <<PY
import json
rows = [[json.loads(l) for l in open("gpurun_out/r04v_gemm_probe_%s.jsonl" % v)] for v in ("keep", "stage")]
for x, y in zip(*rows):
    print(x["case"], "| shipped", round(x["ms"], 4), round(x["tflops_sustained"], 1), "| staged", round(y["ms"], 4), round(y["tflops_sustained"], 1))
PY
tail -3 gpurun_out/r04v_gemm_probe.err
