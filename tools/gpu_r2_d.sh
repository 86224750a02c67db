#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest "tests/test_gpu_ppo_fused.py::test_one_minibatch_matches_fp32_torch" -m gpu -q -s 2>&1 | grep -E "rel errors|AssertionError|passed|failed" | cut -c1-1500 | tee gpurun_out/r2d_fused.txt
timeout 300 python tools/profile_ppo_fused.py 1048576 32768 bf16x3 2>&1 | grep -A40 "launch sequence" | tee gpurun_out/r2d_seq_x3.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_gemm -s 28 -c 14 -f -o gpurun_out/r2d_umma python tools/profile_ppo_fused.py 65536 32768 bf16x3 > gpurun_out/r2d_ncu.log 2>&1; tail -3 gpurun_out/r2d_ncu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:head_kernel -s 2 -c 1 -f -o gpurun_out/r2d_head python tools/profile_ppo_fused.py 65536 32768 bf16x3 > gpurun_out/r2d_ncu2.log 2>&1; tail -3 gpurun_out/r2d_ncu2.log
ls -la gpurun_out/*.ncu-rep
