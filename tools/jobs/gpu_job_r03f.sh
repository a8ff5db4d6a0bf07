#!/bin/bash
# GPU job 3F: programmatic dependent launch on the GEMM / attention / ln_modulate kernels: parity with X2I_PDL=1, then A/B of the bench line.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r03f_build.log 2>&1
X2I_PDL=1 timeout 1200 python -m pytest tests/test_gpu_flux.py tests/test_gpu_kernels.py tests/test_gpu_parity_full.py tests/test_mllm_prefill.py -x -q -m gpu > gpurun_out/r03f_tests_pdl.log 2>&1; echo "tests (PDL) rc=$?" | tee gpurun_out/r03f_rc.log
tail -3 gpurun_out/r03f_tests_pdl.log
for rep in 1 2; do for pdl in 0 1; do
X2I_PDL=$pdl timeout 600 python bench.py --steps 20 --warmup 5 --no-train --no-cpu-baseline --no-library-baseline > gpurun_out/r03f_bench_pdl${pdl}_$rep.json 2> gpurun_out/r03f_bench_pdl${pdl}_$rep.err; echo "pdl=$pdl rep=$rep rc=$?"
python - <<PY
import json
j = json.loads([l for l in open("gpurun_out/r03f_bench_pdl${pdl}_$rep.json") if l.startswith("{")][0])
print("   value", round(j["value"], 3), "ms", round(j["ms_per_step"], 3), "e2e", round(j["e2e"]["value"], 3), "attn ms", round(j["roofline"]["ms_per_launch"], 4), j["clocks"]["sm_mhz"])
PY
done; done
