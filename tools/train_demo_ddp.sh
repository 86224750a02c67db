#!/bin/bash
# Data-parallel PPO training demo: one process per GPU, env shards + NCCL gradient all-reduce.  Run under gpurun --gpus N.
mkdir -p gpurun_out
NG=${1:-2}
timeout ${3:-300} python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29555 \
   -m drl_dronenavigation_b200.simulation_controller --agent PPO --run_type full --num_envs 8192 \
   --total_timesteps 4e9 --rollout_steps 128 --minibatch 16384 --savemodel f --max_seconds ${2:-150} > gpurun_out/train_ppo_ddp$NG.log 2>&1
grep -v "^\*\|OMP" gpurun_out/train_ppo_ddp$NG.log | tail -25
