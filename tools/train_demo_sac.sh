#!/bin/bash
# SAC through the reference's command line on one B200 (circle track, drag + ground effect are bench-only; this is the
# manager path: --agent SAC --run_type full with --savemodel, then --run_type cont from the written archive + replay buffer).
# Run under gpurun.  usage: train_demo_sac.sh NUM_ENVS MAX_SECONDS
mkdir -p gpurun_out
N=${1:-4096}; MS=${2:-60}
rm -rf Sol/model_chkpts/SAC_save_*
timeout $((MS + 120)) python -m drl_dronenavigation_b200.simulation_controller --agent SAC --run_type full --num_envs $N \
   --total_timesteps 4e9 --savemodel t --max_seconds $MS --tensorboard gpurun_out/tb_sac > gpurun_out/train_sac.log 2>&1
grep "steps\|final" gpurun_out/train_sac.log | awk 'NR%4==1' | tail -12; tail -2 gpurun_out/train_sac.log
CHK=$(ls -d Sol/model_chkpts/SAC_save_* | tail -1); ls -la $CHK
timeout 180 python -m drl_dronenavigation_b200.simulation_controller --agent SAC --run_type cont --num_envs $N --total_timesteps 4e9 \
   --savemodel f --max_seconds 15 --model_path $CHK/success_model.zip > gpurun_out/train_sac_cont.log 2>&1
grep "replay buffer\|final" gpurun_out/train_sac_cont.log; tail -2 gpurun_out/train_sac_cont.log
