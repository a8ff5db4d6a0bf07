#!/bin/bash
# round-2 GPU job Z: GroupNorm-backward / column-sum tails, LayerNorm from the fp32 plane, hint-feature cache: tests, LC train + editing step, prefill.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02z_build.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_conv_bwd.py tests/test_controlnext.py tests/test_gpu_bwd_kernels.py tests/test_gpu_kernels.py tests/test_gpu_train.py tests/test_vae.py -x -q -m gpu > gpurun_out/r02z_tests.log 2>&1; echo "tests rc=$?" | tee gpurun_out/r02z_rc.log
tail -4 gpurun_out/r02z_tests.log
timeout 600 python tools/bench_lightcontrol_train.py > gpurun_out/r02z_lc_train.json 2> gpurun_out/r02z_lc_train.err; cat gpurun_out/r02z_lc_train.json; tail -3 gpurun_out/r02z_lc_train.err
timeout 600 python tools/bench_lightcontrol.py > gpurun_out/r02z_lc_edit.json 2> gpurun_out/r02z_lc_edit.err; cat gpurun_out/r02z_lc_edit.json; tail -3 gpurun_out/r02z_lc_edit.err
timeout 120 python tools/probe_projconv.py > gpurun_out/r02z_probe.jsonl 2> gpurun_out/r02z_probe.err; cut -c1-400 gpurun_out/r02z_probe.jsonl
