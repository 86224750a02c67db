#!/bin/bash
# PPO training demo on one B200: reference flags, reference hyper-parameters, GPU environment.  Run under gpurun.
mkdir -p gpurun_out
timeout ${2:-420} python -m drl_dronenavigation_b200.simulation_controller --agent PPO --run_type full --num_envs ${1:-8192} \
   --total_timesteps 4e9 --rollout_steps 128 --minibatch 16384 --savemodel f --max_seconds ${3:-300} \
   --tensorboard gpurun_out/tb_ppo > gpurun_out/train_ppo.log 2>&1
tail -40 gpurun_out/train_ppo.log
