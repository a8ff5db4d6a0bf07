#!/bin/bash
# round-2 GPU job R: wgrad producer without per-k-block divisions + warp-per-channel GroupNorm backward reduction: tests, LC train step.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02r_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_conv_bwd.py tests/test_controlnext.py -x -q -m gpu > gpurun_out/r02r_tests.log 2>&1; echo "tests rc=$?" | tee gpurun_out/r02r_rc.log
tail -4 gpurun_out/r02r_tests.log
timeout 600 python tools/bench_lightcontrol_train.py > gpurun_out/r02r_lc_train.json 2> gpurun_out/r02r_lc_train.err; cat gpurun_out/r02r_lc_train.json; tail -3 gpurun_out/r02r_lc_train.err
