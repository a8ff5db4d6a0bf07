#!/bin/bash
# round-2 GPU job K: tcgen05.mma shape / operand-form rate micro-benchmark (input to the attention-backward redesign)
mkdir -p gpurun_out
timeout 120 tools/ubench/mma_rate 2>&1 | tee gpurun_out/r02k_mma_rate.txt
