#!/bin/bash
# round-2 GPU job I: a12 legacy projectors + everything touched since the last full run (quick), then MLLM regression.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02i_build.log 2>&1
timeout 900 python -m pytest tests/test_proj_legacy.py tests/test_mllm_prefill.py tests/test_vae.py -x -q -m gpu > gpurun_out/r02i_tests.log 2>&1; echo "tests rc=$?" | tee gpurun_out/r02i_rc.log
grep -v "^$" gpurun_out/r02i_tests.log | tail -25
