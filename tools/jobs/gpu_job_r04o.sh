#!/bin/bash
# GPU job 4O: 256-bit loads / stores in the GEMM and attention epilogues: tests, gemm_probe, and the denoise step A/B against the previous
# library (libx2i_old.so = HEAD before this change) on one box.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_flux.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/r04o_tests.log 2>&1; echo "tests rc=$?" | tee gpurun_out/r04o_rc.log
tail -3 gpurun_out/r04o_tests.log
timeout 600 python tools/gemm_probe.py > gpurun_out/r04o_gemm_probe.jsonl 2> gpurun_out/r04o_gemm_probe.err
python - <<PY
import json
for l in open("gpurun_out/r04o_gemm_probe.jsonl"):
    j = json.loads(l)
    print(j["case"], "|", round(j["ms"], 4), "ms", round(j["tflops_sustained"], 1), "TF", j["sm_mhz_median"], "MHz", round(j["tensor_util_at_clock"], 3))
PY
cp x2i_b200/libx2i_b200.so /tmp/new.so
for rep in 1 2; do for which in old new; do
if [ $which = old ]; then cp libx2i_old.so x2i_b200/libx2i_b200.so; else cp /tmp/new.so x2i_b200/libx2i_b200.so; fi
timeout 600 python bench.py --steps 20 --warmup 5 --no-train --no-cpu-baseline --no-library-baseline > gpurun_out/r04o_b.json 2> gpurun_out/r04o_b.err; echo "$which rep=$rep rc=$?"
python - <<PY
import json
j = json.loads([l for l in open("gpurun_out/r04o_b.json") if l.startswith("{")][0])
r = j["roofline"]
print("   value", round(j["value"], 3), "ms", round(j["ms_per_step"], 3), "e2e", round(j["e2e"]["value"], 3), "attn ms", round(r["ms_per_launch"], 4), "burst frac", round(r["frac_of_burst_peak"], 4), "iso", round(r["isolated_tflops"], 1), j["clocks"]["sm_mhz"])
PY
done; done
cp /tmp/new.so x2i_b200/libx2i_b200.so
