#!/bin/bash
# GPU job 3R: the secondary benchmark tools still run on the final tree (numbers for the record).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r03r_build.log 2>&1
for t in bench_512 bench_vae bench_lightcontrol bench_mllm; do
  timeout 600 python tools/$t.py > gpurun_out/r03r_$t.json 2> gpurun_out/r03r_$t.err; echo "$t rc=$?"; cut -c1-600 gpurun_out/r03r_$t.json; tail -1 gpurun_out/r03r_$t.err
done
timeout 900 python tools/bench_train.py --batch 1 > gpurun_out/r03r_bench_train_b1.json 2> gpurun_out/r03r_bench_train_b1.err; echo "bench_train rc=$?"; cut -c1-600 gpurun_out/r03r_bench_train_b1.json; tail -1 gpurun_out/r03r_bench_train_b1.err
