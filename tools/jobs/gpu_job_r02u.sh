#!/bin/bash
# round-2 GPU job U: wave-model tile selection for plain GEMMs: MLLM prefill / projector / legacy tests, prefill timing, bench line.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02u_build.log 2>&1
timeout 900 python -m pytest tests/test_mllm_prefill.py  -x -q -m gpu > gpurun_out/r02u_tests.log 2>&1; echo "tests rc=$?" | tee gpurun_out/r02u_rc.log
tail -3 gpurun_out/r02u_tests.log
timeout 300 python tools/bench_mllm.py > gpurun_out/r02u_bench_mllm.json 2> gpurun_out/r02u_bench_mllm.err; cat gpurun_out/r02u_bench_mllm.json; tail -3 gpurun_out/r02u_bench_mllm.err
