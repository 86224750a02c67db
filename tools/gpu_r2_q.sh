#!/bin/bash
# ncu --set full captures: env step kernel variants + the final PPO contraction kernels + head kernel.
# gpurun copies back at most 64 MiB: every report is exported to CSV on the box and only two reports travel.
mkdir -p gpurun_out
cap() {  # name, then profile_step.py args
  name=$1; shift
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 30 -c 1 -f -o gpurun_out/r2q_$name python tools/profile_step.py "$@" > gpurun_out/r2q_$name.log 2>&1
  ncu -i gpurun_out/r2q_$name.ncu-rep --page raw --csv > gpurun_out/r2q_$name.raw.csv 2>/dev/null
  tail -1 gpurun_out/r2q_$name.log
}
cap n4096_s8 4096 8 34
cap norm_n4096_s8 4096 8 34 --norm-obs
cap norm_n12_s1 12 1 34 --norm-obs
cap phys3_n16384_s8 16384 8 34 --physics gnd_drag
cap full_rw3_n131072_s8 131072 8 34 --reward-id 3
cap full_rw8_n131072_s8 131072 8 34 --reward-id 8
cap full_rw9_n131072_s8 131072 8 34 --reward-id 9
cap n4194304_s8 4194304 8 34
cap n4194304_s1 4194304 1 34
ncu -i gpurun_out/r2q_n4194304_s8.ncu-rep --page source --csv > gpurun_out/r2q_n4194304_s8.source.csv 2>/dev/null
for f in norm_n4096_s8 norm_n12_s1 phys3_n16384_s8 full_rw3_n131072_s8 full_rw8_n131072_s8 full_rw9_n131072_s8 n4194304_s8 n4194304_s1; do rm -f gpurun_out/r2q_$f.ncu-rep; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"umma_gemm|head_kernel|reduce_kernel" -s 25 -c 19 -f -o gpurun_out/r2q_ppo python tools/profile_ppo_fused.py 65536 32768 bf16x3 > gpurun_out/r2q_ppo.log 2>&1; tail -1 gpurun_out/r2q_ppo.log
ncu -i gpurun_out/r2q_ppo.ncu-rep --page raw --csv > gpurun_out/r2q_ppo.raw.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2q_launches.csv python bench.py --steps 64 --warmup 3 --no-cpu --no-vecenv --no-ppo --no-configs --rotating-handles 8 --sweep 4194304 > gpurun_out/r2q_ncu_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2q_launches_ppo.csv -k regex:"dnmma|dnppo" python tools/profile_ppo_fused.py 65536 32768 bf16x3 > /dev/null 2>&1
du -sh gpurun_out; ls gpurun_out
