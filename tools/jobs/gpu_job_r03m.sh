#!/bin/bash
# GPU job 3M: LightControl editing step through the CUDA graph (nets' mid features eager, transformer + injection convs replayed): tests, timing.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r03m_build.log 2>&1
timeout 1200 python -m pytest tests/test_controlnext.py tests/test_gpu_parity_full.py tests/test_gpu_flux.py -x -q -m gpu > gpurun_out/r03m_tests.log 2>&1; echo "tests rc=$?" | tee gpurun_out/r03m_rc.log
tail -4 gpurun_out/r03m_tests.log
timeout 600 python tools/bench_lightcontrol.py > gpurun_out/r03m_lc_edit.json 2> gpurun_out/r03m_lc_edit.err; cat gpurun_out/r03m_lc_edit.json; tail -3 gpurun_out/r03m_lc_edit.err
