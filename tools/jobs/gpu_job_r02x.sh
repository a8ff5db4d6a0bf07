#!/bin/bash
# round-2 GPU job X: tensor-core layer-mixing convolution (deep A ring, 56-column tiles): parity tests incl. projector fwd/bwd, timing, prefill.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02x_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_flux.py tests/test_gpu_fullsize.py tests/test_gpu_train.py tests/test_mllm_prefill.py -x -q -m gpu -k "proj or layer_mixing or prefill or distill or train" > gpurun_out/r02x_tests.log 2>&1; echo "tests rc=$?" | tee gpurun_out/r02x_rc.log
tail -4 gpurun_out/r02x_tests.log
timeout 120 python tools/probe_projconv.py > gpurun_out/r02x_probe.jsonl 2> gpurun_out/r02x_probe.err; cat gpurun_out/r02x_probe.jsonl; tail -3 gpurun_out/r02x_probe.err
timeout 300 python tools/bench_rowwise.py > gpurun_out/r02x_rowwise.jsonl 2> gpurun_out/r02x_rowwise.err; tail -2 gpurun_out/r02x_rowwise.jsonl | cut -c1-300
timeout 300 python tools/bench_mllm.py > gpurun_out/r02x_bench_mllm.json 2> gpurun_out/r02x_bench_mllm.err; cat gpurun_out/r02x_bench_mllm.json
