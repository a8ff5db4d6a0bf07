#!/bin/bash
# round-2 GPU job Y: refresh the ncu evidence after this session's kernel changes.  Never a bench number.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02y_build.log 2>&1
rm -f gpurun_out/launches.csv gpurun_out/prof_*_raw.csv
bash tools/profile_cmds.sh > gpurun_out/profile_cmds.log 2>&1
export X2I_NCU=0
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:"kd_row|proj_conv_tc|ln_rows_f32|proj_mix_ln|skinny_linear_t_final|colsum_final|mmdit_attention_bwd|gn_bwd|ln_modulate|gemm_tcgen05_kernel" -c 40 -o gpurun_out/prof_train -f python tools/profile_bwd_kernels.py > gpurun_out/ncu_prof_train.log 2>&1
ncu -i gpurun_out/prof_train.ncu-rep --page raw --csv > gpurun_out/prof_train_raw.csv 2>/dev/null; rm -f gpurun_out/prof_train.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/mllm_launches.csv python tools/bench_mllm.py --steps 1 --warmup 1 > gpurun_out/ncu_mllm.log 2>&1
timeout 300 python tools/bench_rowwise.py > gpurun_out/r02y_rowwise.jsonl 2> gpurun_out/r02y_rowwise.err
tail -2 gpurun_out/ncu_prof_train.log; wc -l gpurun_out/prof_*_raw.csv gpurun_out/launches.csv gpurun_out/mllm_launches.csv; cat gpurun_out/r02y_rowwise.jsonl | cut -c1-250
