#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mlp.py -m gpu -q -x 2>&1 | grep -vE "^\s*$" | cut -c1-300 | tail -30 | tee gpurun_out/r2e_mlp.txt
timeout 900 python -m pytest tests/test_gpu_ppo_fused.py -m gpu -q -x 2>&1 | grep -vE "^\s*$" | cut -c1-300 | tail -30 | tee gpurun_out/r2e_fused.txt
timeout 300 python tools/profile_ppo_fused.py 1048576 32768 bf16x3 2>&1 | grep -v Warn | grep -B2 -A48 "^update" | tee gpurun_out/r2e_prof_x3.txt
