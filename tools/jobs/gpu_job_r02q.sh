#!/bin/bash
# round-2 GPU job Q: ncu launch list of one LightControl train step (implicit wgrad build).  Never a bench number.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02q_build.log 2>&1
X2I_NCU=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lc_train_launches.csv \
  python tools/bench_lightcontrol_train.py --steps 1 --warmup 1 > gpurun_out/r02q_ncu.log 2>&1
tail -2 gpurun_out/r02q_ncu.log; wc -l gpurun_out/lc_train_launches.csv
