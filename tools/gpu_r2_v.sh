#!/bin/bash
# ncu --set full of one minibatch of the grouped PPO update (10 kernels), exported to CSV on the box
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"umma_gemm|head_kernel|reduce_kernel" -s 20 -c 10 -f -o gpurun_out/r2v_ppo python tools/profile_ppo_fused.py 65536 32768 bf16x3 > gpurun_out/r2v_ppo.log 2>&1; tail -2 gpurun_out/r2v_ppo.log
ncu -i gpurun_out/r2v_ppo.ncu-rep --page raw --csv > gpurun_out/r2v_ppo.raw.csv 2>/dev/null
rm -f gpurun_out/r2v_ppo.ncu-rep
ls -la gpurun_out
