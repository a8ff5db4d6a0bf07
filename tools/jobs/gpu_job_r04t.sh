#!/bin/bash
# GPU job 4T: epilogue inputs (residual + gate, RoPE entries) requested together with the accumulators instead of after them: tests,
# gemm_probe against the previous library (libx2i_old.so = HEAD), denoise step A/B.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_flux.py tests/test_mllm_prefill.py -x -q -m gpu > gpurun_out/r04t_tests.log 2>&1; echo "tests rc=$?" | tee gpurun_out/r04t_rc.log
tail -3 gpurun_out/r04t_tests.log
cp x2i_b200/libx2i_b200.so /tmp/new.so
for v in old new; do
if [ $v = old ]; then cp libx2i_old.so x2i_b200/libx2i_b200.so; else cp /tmp/new.so x2i_b200/libx2i_b200.so; fi
timeout 600 python tools/gemm_probe.py > gpurun_out/r04t_gemm_probe_$v.jsonl 2>> gpurun_out/r04t_gemm_probe.err
done
python - <<PY
import json
rows = [[json.loads(l) for l in open("gpurun_out/r04t_gemm_probe_%s.jsonl" % v)] for v in ("old", "new")]
for x, y in zip(*rows):
    print(x["case"], "| old", round(x["ms"], 4), round(x["tflops_sustained"], 1), "| new", round(y["ms"], 4), round(y["tflops_sustained"], 1))
PY
for which in old new; do
if [ $which = old ]; then cp libx2i_old.so x2i_b200/libx2i_b200.so; else cp /tmp/new.so x2i_b200/libx2i_b200.so; fi
timeout 600 python bench.py --steps 20 --warmup 5 --no-train --no-cpu-baseline --no-library-baseline > gpurun_out/r04t_b.json 2> gpurun_out/r04t_b.err; echo "$which rc=$?"
python - <<PY
import json
j = json.loads([l for l in open("gpurun_out/r04t_b.json") if l.startswith("{")][0])
r = j["roofline"]
print("   value", round(j["value"], 3), "ms", round(j["ms_per_step"], 3), "e2e", round(j["e2e"]["value"], 3), "attn ms", round(r["ms_per_launch"], 4), j["clocks"]["sm_mhz"])
PY
done
cp /tmp/new.so x2i_b200/libx2i_b200.so
