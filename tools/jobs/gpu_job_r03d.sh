#!/bin/bash
# GPU job 3D (2 GPUs): full validation of the current tree: whole -m gpu suite, smoke, LightControl trainer under torchrun (DP + side streams), bench N=1.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r03d_build.log 2>&1
( time timeout 2400 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/r03d_pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" | tee gpurun_out/r03d_rc.log
tail -4 gpurun_out/r03d_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r03d_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/r03d_rc.log; tail -1 gpurun_out/r03d_smoke.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/bench_lightcontrol_train.py > gpurun_out/r03d_lc_train_n2.json 2> gpurun_out/r03d_lc_train_n2.err; echo "lc train n2 rc=$?" | tee -a gpurun_out/r03d_rc.log; cat gpurun_out/r03d_lc_train_n2.json; tail -2 gpurun_out/r03d_lc_train_n2.err
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r03d_bench_n1.json 2> gpurun_out/r03d_bench_n1.err; echo "bench rc=$?" | tee -a gpurun_out/r03d_rc.log
head -c 600 gpurun_out/r03d_bench_n1.json; tail -4 gpurun_out/r03d_bench_n1.err
