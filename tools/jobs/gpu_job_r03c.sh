#!/bin/bash
# GPU job 3C: control nets on side streams in the LightControl trainer: gradient tests, then A/B of the train step on one box.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r03c_build.log 2>&1
timeout 900 python -m pytest tests/test_controlnext.py -x -q -m gpu > gpurun_out/r03c_tests.log 2>&1; echo "tests rc=$?" | tee gpurun_out/r03c_rc.log
tail -5 gpurun_out/r03c_tests.log
for s in 1 4 8 2; do
X2I_CN_STREAMS=$s timeout 600 python tools/bench_lightcontrol_train.py > gpurun_out/r03c_lc_train_s$s.json 2> gpurun_out/r03c_lc_train_s$s.err; echo "streams=$s"; cat gpurun_out/r03c_lc_train_s$s.json; tail -2 gpurun_out/r03c_lc_train_s$s.err
done
