#!/bin/bash
mkdir -p gpurun_out
timeout 2700 python -m pytest tests -m gpu -q 2>&1 | grep -vE "^\s*$" | cut -c1-300 | tail -40 | tee gpurun_out/r2n_pytest_gpu.txt
timeout 300 python tools/profile_ppo_fused.py 1048576 32768 bf16x3 2>&1 | grep -v Warn | grep -A40 "^update" | grep -v "at::\|at_cuda\|Mem" | tee gpurun_out/r2n_prof_x3.txt
