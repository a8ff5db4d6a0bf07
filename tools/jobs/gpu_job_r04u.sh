#!/bin/bash
# GPU job 4U (2 GPUs): the driver's launch line at N=2 with the final tree (bench incl. the DP distillation step), reference arm at N=2.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r04u_build.log 2>&1
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r04u_bench_n2.json 2> gpurun_out/r04u_bench_n2.err; echo "bench n2 rc=$?" | tee gpurun_out/r04u_rc.log
python - <<PY
import json
j = json.loads([l for l in open("gpurun_out/r04u_bench_n2.json") if l.startswith("{")][0])
r = j["roofline"]
print("value", round(j["value"], 3), "e2e", round(j["e2e"]["value"], 3), "attn ms", round(r["ms_per_launch"], 4), j["clocks"], "distill", j["distill_train"]["samples_per_s"], j["distill_train"]["collective"])
PY
tail -3 gpurun_out/r04u_bench_n2.err
