#!/bin/bash
# round-2 GPU job A: new parity tests + changed tests, bench N=1, compute-sanitizer logs.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/parity_table.jsonl gpurun_out/parity_depth_curve.json
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02a_build.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_parity_full.py -x -q -m gpu -s > gpurun_out/r02a_parity_full.log 2>&1; echo "parity_full rc=$?" | tee -a gpurun_out/r02a_rc.log
timeout 1500 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_parity_full.py > gpurun_out/r02a_pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" | tee -a gpurun_out/r02a_rc.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench_n1.json 2> gpurun_out/r02a_bench_n1.err; echo "bench rc=$?" | tee -a gpurun_out/r02a_rc.log
for tool in synccheck racecheck memcheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/sanitize_small.py > gpurun_out/r02a_sanitizer_$tool.log 2>&1; echo "sanitizer $tool rc=$?" | tee -a gpurun_out/r02a_rc.log
done
tail -3 gpurun_out/r02a_parity_full.log; tail -3 gpurun_out/r02a_pytest_gpu.log; head -c 1500 gpurun_out/r02a_bench_n1.json
