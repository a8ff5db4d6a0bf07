#!/bin/bash
# GPU job 4F: lagged soft-max steps as the default of the persistent attention forward: whole -m gpu suite, smoke, sanitizer (incl. the redo
# pass), A/B of the denoise step (X2I_ATTN_LAG=0/1), default bench line.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r04f_build.log 2>&1
( time timeout 2400 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/r04f_pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" | tee gpurun_out/r04f_rc.log
tail -4 gpurun_out/r04f_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r04f_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/r04f_rc.log; tail -1 gpurun_out/r04f_smoke.log
for tool in synccheck memcheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/sanitize_small.py > gpurun_out/r04f_sanitizer_$tool.log 2>&1; echo "sanitizer $tool rc=$?" | tee -a gpurun_out/r04f_rc.log
  grep "redo pass" gpurun_out/r04f_sanitizer_$tool.log; tail -1 gpurun_out/r04f_sanitizer_$tool.log
done
for rep in 1 2; do for lag in 0 1; do
X2I_ATTN_LAG=$lag timeout 600 python bench.py --steps 20 --warmup 5 --no-train --no-cpu-baseline --no-library-baseline > gpurun_out/r04f_b.json 2> gpurun_out/r04f_b.err; echo "lag=$lag rep=$rep rc=$?"
python - <<PY
import json
j = json.loads([l for l in open("gpurun_out/r04f_b.json") if l.startswith("{")][0])
r = j["roofline"]
print("   value", round(j["value"], 3), "ms", round(j["ms_per_step"], 3), "e2e", round(j["e2e"]["value"], 3), "attn ms", round(r["ms_per_launch"], 4), "burst frac", round(r["frac_of_burst_peak"], 4), "iso", round(r["isolated_tflops"], 1), "lib", round(r["library_tflops"], 1), j["clocks"]["sm_mhz"])
PY
done; done
( time timeout 900 python bench.py ) > gpurun_out/r04f_bench_default.json 2> gpurun_out/r04f_bench_default.err; echo "bench rc=$?" | tee -a gpurun_out/r04f_rc.log
python - <<PY
import json
j = json.loads([l for l in open("gpurun_out/r04f_bench_default.json") if l.startswith("{")][0])
r = j["roofline"]
print("   value", round(j["value"], 3), "ms", round(j["ms_per_step"], 3), "e2e", round(j["e2e"]["value"], 3), "attn ms", round(r["ms_per_launch"], 4), round(r["frac_of_burst_peak"], 4), j["clocks"], "launches", j["gpu_launches"])
print("   distill", round(j["distill_train"]["ms_per_step"], 1), "prefill", round(j["mllm_prefill"]["ms_per_prompt"], 3), "lc", round(j["lightcontrol_train"]["ms_per_step"], 1), "vae", round(j["vae_decode"]["ms_per_decode"], 2), "lib", j["gpu_library_baseline"]["value"], "cpu", j["cpu_baseline"]["value"])
PY
