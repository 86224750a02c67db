"""Where does the PPO update spend its time?  usage: profile_ppo.py [ENVS] [ROLLOUT]   (run under gpurun)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from drl_dronenavigation_b200.ppo import PPOConfig, PPOTrainer

N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = torch.device("cuda", 0)


class A:
    envs, substeps, track, actions = N, 8, "reaching", "saturating"


env = bench.make_env(N, A, dev)
cfg = PPOConfig(n_steps=T, batch_size=max(512, (T * N) // 32), n_epochs=2)
tr = PPOTrainer(env, cfg, rollout_steps=T)
tr.train_iteration()
torch.cuda.synchronize()
t0 = time.perf_counter()
out = tr.train_iteration()
torch.cuda.synchronize()
print("eager iteration", time.perf_counter() - t0, {k: out[k] for k in ("rollout_s", "update_s", "minibatches")})
from torch.profiler import ProfilerActivity, profile
adv, ret = tr.collect_rollouts()
D = env.obs_dim
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    tr.learner.update(tr.b_obs.reshape(-1, D), tr.b_act.reshape(-1, 4), tr.b_logp.reshape(-1), tr.b_val.reshape(-1),
                      adv.reshape(-1), ret.reshape(-1), generator=tr.gen)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    tr.collect_rollouts()
    torch.cuda.synchronize()
    print("collect_rollouts", rep, time.perf_counter() - t0)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    tr.collect_rollouts()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=10, max_name_column_width=60))
