#!/bin/bash
# round-2 GPU job H: persistent KD kernel + proj_mix_ln<R> correctness, then the row-wise kernel bandwidth table.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02h_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_flux.py tests/test_gpu_train.py tests/test_gpu_fullsize.py -x -q -m gpu -k "kd or projector or proj or distill or rowwise" > gpurun_out/r02h_tests.log 2>&1; echo "tests rc=$?" | tee gpurun_out/r02h_rc.log
tail -3 gpurun_out/r02h_tests.log
timeout 300 python tools/bench_rowwise.py > gpurun_out/r02h_rowwise.jsonl 2> gpurun_out/r02h_rowwise.err; cat gpurun_out/r02h_rowwise.jsonl; tail -3 gpurun_out/r02h_rowwise.err
