#!/bin/bash
# GPU job 3L (8 GPUs): the driver's launch line at N=8 with the final tree (bench incl. the DP distillation step).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r03l_build.log 2>&1
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 ) > gpurun_out/r03l_bench_n8.json 2> gpurun_out/r03l_bench_n8.err; echo "bench n8 rc=$?" | tee gpurun_out/r03l_rc.log
python - <<PY
import json
j = json.loads([l for l in open("gpurun_out/r03l_bench_n8.json") if l.startswith("{")][0])
r = j["roofline"]
print("value", round(j["value"], 3), "e2e", round(j["e2e"]["value"], 3), "attn ms", round(r["ms_per_launch"], 4), j["clocks"], "distill", j["distill_train"]["samples_per_s"], j["distill_train"]["collective"])
PY
tail -3 gpurun_out/r03l_bench_n8.err
