#!/bin/bash
# round-2 GPU job S: full validation of the current tree: whole -m gpu suite (as the driver runs it), smoke, bench N=1, sanitizer on the changed backward.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02s_build.log 2>&1
( time timeout 2400 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/r02s_pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" | tee gpurun_out/r02s_rc.log
tail -4 gpurun_out/r02s_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02s_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/r02s_rc.log; tail -2 gpurun_out/r02s_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/r02s_bench_n1.json 2> gpurun_out/r02s_bench_n1.err; echo "bench rc=$?" | tee -a gpurun_out/r02s_rc.log
head -c 2500 gpurun_out/r02s_bench_n1.json; tail -5 gpurun_out/r02s_bench_n1.err
for tool in synccheck memcheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/sanitize_small.py > gpurun_out/r02s_sanitizer_$tool.log 2>&1; echo "sanitizer $tool rc=$?" | tee -a gpurun_out/r02s_rc.log
  tail -2 gpurun_out/r02s_sanitizer_$tool.log
done
