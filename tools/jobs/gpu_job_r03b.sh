#!/bin/bash
# GPU job 3B (2 GPUs): the driver's launch lines at N=2 and N=1 with the current tree (bench incl. the DP distillation step, reference arm).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r03b_build.log 2>&1
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r03b_bench_n2.json 2> gpurun_out/r03b_bench_n2.err; echo "bench n2 rc=$?" | tee gpurun_out/r03b_rc.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r03b_bench_n1.json 2> gpurun_out/r03b_bench_n1.err; echo "bench n1 rc=$?" | tee -a gpurun_out/r03b_rc.log
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 ) > gpurun_out/r03b_ref_n1.json 2> gpurun_out/r03b_ref_n1.err; echo "ref n1 rc=$?" | tee -a gpurun_out/r03b_rc.log
head -c 700 gpurun_out/r03b_bench_n2.json; echo; tail -4 gpurun_out/r03b_bench_n2.err; head -c 500 gpurun_out/r03b_bench_n1.json; echo; tail -4 gpurun_out/r03b_bench_n1.err; head -c 400 gpurun_out/r03b_ref_n1.json; tail -4 gpurun_out/r03b_ref_n1.err
