#!/bin/bash
# PPO training demo on one B200: reference flags, reference hyper-parameters, GPU environment.  Run under gpurun.
# usage: train_demo.sh NUM_ENVS TIMEOUT MAX_SECONDS [extra flags]
mkdir -p gpurun_out
N=${1:-8192}; TO=${2:-420}; MS=${3:-300}; shift 3
timeout $TO python -m drl_dronenavigation_b200.simulation_controller --agent PPO --run_type full --num_envs $N \
   --total_timesteps 4e9 --rollout_steps 128 --minibatch 16384 --savemodel f --max_seconds $MS \
   --tensorboard gpurun_out/tb_ppo "$@" > gpurun_out/train_ppo.log 2>&1
grep "steps\|final" gpurun_out/train_ppo.log | awk 'NR%4==1' | tail -22; tail -2 gpurun_out/train_ppo.log
