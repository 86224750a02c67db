#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ppo_fused.py -m gpu -q -s 2>&1 | grep -vE "^\s*$" | tail -60 > gpurun_out/r2c_fused.txt; cat gpurun_out/r2c_fused.txt
timeout 300 python tools/profile_ppo_fused.py 1048576 32768 bf16x3 2>&1 | tail -32 | tee gpurun_out/r2c_prof_x3.txt
timeout 300 python tools/profile_ppo_fused.py 1048576 32768 bf16 2>&1 | tail -32 | tee gpurun_out/r2c_prof_bf16.txt
