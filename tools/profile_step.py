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
ap.add_argument("--norm-obs", action="store_true", help="fused NormalizeObservation (step_kernel<*, true, *, *>)")
ap.add_argument("--physics", default="dyn", choices=["dyn", "gnd_drag"], help="gnd_drag: step_kernel<3, ...> (BASELINE config 4)")
ap.add_argument("--reward-id", type=int, default=0, help=">= 3: the FULL instantiations (BASELINE config 5)")
ap.add_argument("--act", default="thrust", choices=["thrust", "pid"], help="pid: DSLPIDControl fused into the step (FULL)")
a = ap.parse_args()
dev = torch.device("cuda", 0)
from drl_dronenavigation_b200 import Physics
from drl_dronenavigation_b200.enums import ActionType
kw = {}
if a.norm_obs:
    kw["normalize_obs"] = True
if a.physics == "gnd_drag":
    kw["physics"] = Physics.PYB_GND_DRAG_DW
if a.reward_id:
    kw["reward_id"] = a.reward_id
if a.act == "pid":
    kw["act"] = ActionType.PID
env = bench.make_env(a.envs, a, dev, **kw)
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
