#!/bin/bash
# round-2 GPU job F (2 GPUs): bench.py under torchrun at N=2 (the driver's launch line) incl. the DP distillation step, then N=1 on GPU 0.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02f_build.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02f_bench_n2.json 2> gpurun_out/r02f_bench_n2.err; echo "bench n2 rc=$?" | tee gpurun_out/r02f_rc.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err; echo "bench n1 rc=$?" | tee -a gpurun_out/r02f_rc.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02f_ref_n2.json 2> gpurun_out/r02f_ref_n2.err; echo "ref n2 rc=$?" | tee -a gpurun_out/r02f_rc.log
head -c 600 gpurun_out/r02f_bench_n2.json; echo; tail -3 gpurun_out/r02f_bench_n2.err; head -c 600 gpurun_out/r02f_bench_n1.json; echo; tail -3 gpurun_out/r02f_bench_n1.err; head -c 400 gpurun_out/r02f_ref_n2.json
