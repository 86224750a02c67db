#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ppo_fused.py tests/test_gpu_mlp.py -m gpu -q -s 2>&1 | grep -vE "^\s*$" | cut -c1-400 | tail -40 > gpurun_out/r2c_fused.txt; cat gpurun_out/r2c_fused.txt
timeout 300 python tools/profile_ppo_fused.py 1048576 32768 bf16x3 2>&1 | grep -v Warn | head -16 | tee gpurun_out/r2c_prof_x3.txt
timeout 300 python tools/profile_ppo_fused.py 1048576 32768 bf16 2>&1 | grep -v Warn | head -14 | tee gpurun_out/r2c_prof_bf16.txt
DN_MLP_BN=128 timeout 300 python tools/profile_ppo_fused.py 1048576 32768 bf16x3 2>&1 | grep -v Warn | head -12 | tee gpurun_out/r2c_prof_x3_bn128.txt
