#!/bin/bash
# GPU job 3E: bench line with the new extra objects (default flags, as the driver runs it), 7B-width prefill test.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r03e_build.log 2>&1
timeout 600 python -m pytest tests/test_mllm_prefill.py -x -q -m gpu > gpurun_out/r03e_tests.log 2>&1; echo "tests rc=$?" | tee gpurun_out/r03e_rc.log; tail -3 gpurun_out/r03e_tests.log
( time timeout 900 python bench.py ) > gpurun_out/r03e_bench_default.json 2> gpurun_out/r03e_bench_default.err; echo "bench default rc=$?" | tee -a gpurun_out/r03e_rc.log
tail -5 gpurun_out/r03e_bench_default.err
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r03e_bench_n1.json 2> gpurun_out/r03e_bench_n1.err; echo "bench rc=$?" | tee -a gpurun_out/r03e_rc.log
python - <<'PY'
import json
for f in ("gpurun_out/r03e_bench_default.json", "gpurun_out/r03e_bench_n1.json"):
    j = json.loads([l for l in open(f) if l.startswith("{")][0])
    print(f, round(j["value"], 2), round(j["e2e"]["value"], 2), round(j["roofline"]["ms_per_launch"], 4), round(j["roofline"]["frac_of_burst_peak"], 4), j["clocks"])
    print("  ", j.get("mllm_prefill"), j.get("lightcontrol_train"), j["distill_train"]["ms_per_step"], j["vae_decode"]["ms_per_decode"])
PY
