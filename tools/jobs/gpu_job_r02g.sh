#!/bin/bash
# round-2 GPU job G: ncu evidence for the bench command (launch list + --set full of the attention kernel, GEMMs, row-wise kernels),
# the train step launch list + full captures of the KD / backward kernels, and the MLLM prefill launch list.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02g_build.log 2>&1
bash tools/profile_cmds.sh no-vae > gpurun_out/profile_cmds.log 2>&1
export X2I_NCU=0
timeout 900 ncu --set full --clock-control none -k regex:"kd_row|proj_mix_ln|skinny_linear_t_final|colsum_final|mmdit_attention_bwd" -c 14 -o gpurun_out/prof_train -f python tools/profile_bwd_kernels.py > gpurun_out/ncu_prof_train.log 2>&1
ncu -i gpurun_out/prof_train.ncu-rep --page raw --csv > gpurun_out/prof_train_raw.csv 2>/dev/null; rm -f gpurun_out/prof_train.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/mllm_launches.csv python tools/bench_mllm.py --steps 1 --warmup 1 > gpurun_out/ncu_mllm.log 2>&1
timeout 300 python tools/bench_mllm.py --steps 5 --warmup 2 > gpurun_out/r02g_bench_mllm.json 2> gpurun_out/r02g_bench_mllm.err
cat gpurun_out/r02g_bench_mllm.json; du -sh gpurun_out; ls gpurun_out | tail -30
