#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_ppo_fused.py -m gpu -q -x 2>&1 | grep -vE "^\s*$" | cut -c1-300 | tail -12 | tee gpurun_out/r2k_fused.txt
timeout 300 python tools/profile_ppo_fused.py 1048576 32768 bf16x3 2>&1 | grep -v Warn | grep -A40 "^update" | grep -v "at::\|at_cuda\|Mem" | tee gpurun_out/r2k_prof_x3.txt
timeout 300 python tools/profile_ppo_fused.py 1048576 32768 bf16 2>&1 | grep -v Warn | grep -A12 "^update" | grep -v "at::\|at_cuda\|Mem" | tee gpurun_out/r2k_prof_bf16.txt
