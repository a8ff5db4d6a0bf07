#!/bin/bash
# GPU job 3K (2 GPUs): final evidence with the final tree: ncu launch list + attention --set full capture of the bench command, bench lines N=2 and N=1.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r03k_build.log 2>&1
rm -f gpurun_out/launches.csv gpurun_out/prof_attn_raw.csv
CUDA_VISIBLE_DEVICES=0 bash tools/profile_cmds.sh no-gemm > gpurun_out/profile_cmds.log 2>&1
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r03k_bench_n2.json 2> gpurun_out/r03k_bench_n2.err; echo "bench n2 rc=$?" | tee gpurun_out/r03k_rc.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r03k_bench_n1.json 2> gpurun_out/r03k_bench_n1.err; echo "bench n1 rc=$?" | tee -a gpurun_out/r03k_rc.log
python - <<PY
import json
for f in ("gpurun_out/r03k_bench_n2.json", "gpurun_out/r03k_bench_n1.json"):
    j = json.loads([l for l in open(f) if l.startswith("{")][0])
    r = j["roofline"]
    print(f, "value", round(j["value"], 3), "e2e", round(j["e2e"]["value"], 3), "attn ms", round(r["ms_per_launch"], 4), round(r["frac_of_burst_peak"], 4), j["clocks"]["sm_mhz"], "distill", round(j["distill_train"]["samples_per_s"], 2), j["distill_train"]["collective"]["ms"])
PY
wc -l gpurun_out/launches.csv gpurun_out/prof_attn_raw.csv
