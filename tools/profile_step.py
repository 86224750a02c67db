"""Tiny driver for ncu: builds one env, runs a few fused steps.  usage: profile_step.py ENVS SUBSTEPS [STEPS] [TRACK]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse

import torch

import bench

ap = argparse.ArgumentParser()
ap.add_argument("envs", type=int)
ap.add_argument("substeps", type=int)
ap.add_argument("steps", type=int, nargs="?", default=6)
ap.add_argument("track", nargs="?", default="circle")
ap.add_argument("--actions", default="saturating")
ap.add_argument("--many", type=int, default=0, help="use dn_step_many with this many steps per launch")
a = ap.parse_args()
dev = torch.device("cuda", 0)
env = bench.make_env(a.envs, a, dev)
env.reset()
acts = bench.make_actions(4, a.envs, a.actions, dev, seed=1)
if a.many:
    big = bench.make_actions(a.many, a.envs, a.actions, dev, seed=2)
    for _ in range(a.steps):
        env.step_many(big, per_step_outputs=False)
else:
    for k in range(a.steps):
        env.step(acts[k % 4])
torch.cuda.synchronize()
print("ok", env.episode_stats())
