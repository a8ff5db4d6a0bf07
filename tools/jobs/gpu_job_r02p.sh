#!/bin/bash
# round-2 GPU job P: implicit conv weight gradient + hoisted modulation: parity tests, then the LightControl train step and the bench line.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02p_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_conv_bwd.py tests/test_controlnext.py tests/test_gpu_flux.py tests/test_vae.py -x -q -m gpu > gpurun_out/r02p_tests.log 2>&1; echo "tests rc=$?" | tee gpurun_out/r02p_rc.log
tail -15 gpurun_out/r02p_tests.log
timeout 600 python tools/bench_lightcontrol_train.py > gpurun_out/r02p_lc_train.json 2> gpurun_out/r02p_lc_train.err; cat gpurun_out/r02p_lc_train.json; tail -3 gpurun_out/r02p_lc_train.err
