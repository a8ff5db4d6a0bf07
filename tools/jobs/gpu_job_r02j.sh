#!/bin/bash
# round-2 GPU job J: narrow-matrix column sums -- backward tests, then the LightControl train step timing.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02j_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_bwd_kernels.py tests/test_gpu_conv_bwd.py tests/test_controlnext.py tests/test_gpu_train.py -x -q -m gpu > gpurun_out/r02j_tests.log 2>&1; echo "tests rc=$?" | tee gpurun_out/r02j_rc.log
tail -3 gpurun_out/r02j_tests.log
timeout 600 python tools/bench_lightcontrol_train.py > gpurun_out/r02j_lc_train.json 2> gpurun_out/r02j_lc_train.err; cat gpurun_out/r02j_lc_train.json; tail -3 gpurun_out/r02j_lc_train.err
