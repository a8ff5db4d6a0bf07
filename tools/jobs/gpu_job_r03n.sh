#!/bin/bash
# GPU job 3N: final validation of the final tree: whole -m gpu suite, smoke, default bench line (as the driver runs it), sanitizer.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r03n_build.log 2>&1
( time timeout 2400 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/r03n_pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" | tee gpurun_out/r03n_rc.log
tail -4 gpurun_out/r03n_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r03n_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/r03n_rc.log; tail -1 gpurun_out/r03n_smoke.log
for tool in synccheck memcheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/sanitize_small.py > gpurun_out/r03n_sanitizer_$tool.log 2>&1; echo "sanitizer $tool rc=$?" | tee -a gpurun_out/r03n_rc.log
  tail -1 gpurun_out/r03n_sanitizer_$tool.log
done
( time timeout 900 python bench.py ) > gpurun_out/r03n_bench_default.json 2> gpurun_out/r03n_bench_default.err; echo "bench rc=$?" | tee -a gpurun_out/r03n_rc.log
python - <<PY
import json
j = json.loads([l for l in open("gpurun_out/r03n_bench_default.json") if l.startswith("{")][0])
r = j["roofline"]
print("   value", round(j["value"], 3), "ms", round(j["ms_per_step"], 3), "e2e", round(j["e2e"]["value"], 3), "attn ms", round(r["ms_per_launch"], 4), round(r["frac_of_burst_peak"], 4), j["clocks"], "launches", j["gpu_launches"])
print("   distill", round(j["distill_train"]["ms_per_step"], 1), "prefill", round(j["mllm_prefill"]["ms_per_prompt"], 3), "lc", round(j["lightcontrol_train"]["ms_per_step"], 1), "vae", round(j["vae_decode"]["ms_per_decode"], 2), "lib", j["gpu_library_baseline"]["value"], "cpu", j["cpu_baseline"]["value"])
PY
