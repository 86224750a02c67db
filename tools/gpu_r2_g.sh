#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:umma_gemm -s 22 -c 16 -f -o gpurun_out/r2g_umma python tools/profile_ppo_fused.py 65536 32768 bf16x3 > gpurun_out/r2g_ncu.log 2>&1; tail -2 gpurun_out/r2g_ncu.log
ls -la gpurun_out/r2g_umma.ncu-rep
