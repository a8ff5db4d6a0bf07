#!/bin/bash
# GPU job 3H: PDL on by default (GEMM, attention fwd / bwd, ln_modulate): whole -m gpu suite, sanitizer, train-step and bench numbers.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r03h_build.log 2>&1
( time timeout 2400 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/r03h_pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" | tee gpurun_out/r03h_rc.log
tail -4 gpurun_out/r03h_pytest_gpu.log
for tool in synccheck memcheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/sanitize_small.py > gpurun_out/r03h_sanitizer_$tool.log 2>&1; echo "sanitizer $tool rc=$?" | tee -a gpurun_out/r03h_rc.log
  tail -1 gpurun_out/r03h_sanitizer_$tool.log
done
for pdl in 0 1; do
X2I_PDL=$pdl timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > gpurun_out/r03h_bench_pdl$pdl.json 2> gpurun_out/r03h_bench_pdl$pdl.err; echo "bench pdl=$pdl rc=$?" | tee -a gpurun_out/r03h_rc.log
python - <<PY
import json
j = json.loads([l for l in open("gpurun_out/r03h_bench_pdl$pdl.json") if l.startswith("{")][0])
r = j["roofline"]
print("   value", round(j["value"], 3), "ms", round(j["ms_per_step"], 3), "e2e", round(j["e2e"]["value"], 3), "attn ms", round(r["ms_per_launch"], 4), j["clocks"]["sm_mhz"])
print("   distill", round(j["distill_train"]["ms_per_step"], 1), "prefill", round(j["mllm_prefill"]["ms_per_prompt"], 3), "lc", round(j["lightcontrol_train"]["ms_per_step"], 1), "vae", round(j["vae_decode"]["ms_per_decode"], 2))
PY
done
