#!/bin/bash
# round-2 GPU job C: re-run the tests fixed after job A; soft-max POLY8 sweep + clock64 trace of the single-CTA attention kernel under
# sustained load (experiments build, rebuilt to the production library afterwards on the box only).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02c_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity_full.py -x -q -m gpu -s -k "config2 or config5 or real_width" > gpurun_out/r02c_parity_rest.log 2>&1; echo "parity rest rc=$?" | tee gpurun_out/r02c_rc.log
timeout 900 python -m pytest tests/test_controlnext.py tests/test_gpu_train.py tests/test_vae.py tests/test_gpu_flux.py -q -m gpu > gpurun_out/r02c_pytest_changed.log 2>&1; echo "changed tests rc=$?" | tee -a gpurun_out/r02c_rc.log
tail -4 gpurun_out/r02c_parity_rest.log; tail -6 gpurun_out/r02c_pytest_changed.log
X2I_BUILD_EXPERIMENTS=1 python -c "import __graft_entry__ as g; g.build(force=True)" >> gpurun_out/r02c_build.log 2>&1
: > gpurun_out/r02c_poly_sweep.jsonl
for p8 in 0 1 2 3 4; do
  X2I_ATTN_PAIR=0 X2I_ATTN_POLY8=$p8 timeout 120 python tools/attn_probe.py --tag "single poly8=$p8" >> gpurun_out/r02c_poly_sweep.jsonl 2>> gpurun_out/r02c_probe.err
done
timeout 120 python tools/attn_probe.py --sdpa >> gpurun_out/r02c_poly_sweep.jsonl 2>> gpurun_out/r02c_probe.err
X2I_ATTN_PAIR=0 X2I_ATTN_DBG=1 timeout 120 python tools/attn_probe.py --seconds 0.2 --tag trace > gpurun_out/r02c_trace.out 2> gpurun_out/r02c_trace.err
cat gpurun_out/r02c_poly_sweep.jsonl; grep ATTTRACE gpurun_out/r02c_trace.err | head -20; tail -3 gpurun_out/r02c_probe.err
