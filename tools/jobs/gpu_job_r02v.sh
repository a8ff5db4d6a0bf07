#!/bin/bash
# round-2 GPU job V: ln_modulate with the row kept as packed bf16 (64 registers, one wave): parity, bandwidth table, bench line.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02v_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_flux.py tests/test_gpu_bwd_kernels.py -x -q -m gpu > gpurun_out/r02v_tests.log 2>&1; echo "tests rc=$?" | tee gpurun_out/r02v_rc.log
tail -3 gpurun_out/r02v_tests.log
timeout 300 python tools/bench_rowwise.py > gpurun_out/r02v_rowwise.jsonl 2> gpurun_out/r02v_rowwise.err; cat gpurun_out/r02v_rowwise.jsonl | cut -c1-300; tail -2 gpurun_out/r02v_rowwise.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02v_bench_n1.json 2> gpurun_out/r02v_bench_n1.err; echo "bench rc=$?" | tee -a gpurun_out/r02v_rc.log
head -c 900 gpurun_out/r02v_bench_n1.json
