#!/bin/bash
# training demos on the reference's wrapper stack (NormalizeObservation on) with the fused PPO update
mkdir -p gpurun_out
bash tools/train_demo.sh 8192 420 300
cp gpurun_out/train_ppo.log gpurun_out/r2p_train_ppo_norm.log
bash tools/train_demo.sh 8192 240 120 --no_norm_obs t
cp gpurun_out/train_ppo.log gpurun_out/r2p_train_ppo_raw.log
bash tools/train_demo_sac.sh 4096 60
cp gpurun_out/train_sac.log gpurun_out/r2p_train_sac.log; cp gpurun_out/train_sac_cont.log gpurun_out/r2p_train_sac_cont.log
