import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.test_gpu_parity import _make, _actions
envA, _ = _make("circle", 1000, 8)
envB, _ = _make("circle", 1000, 8)
envA.reset(); envB.reset()
T = 40
acts = torch.from_numpy(_actions("saturating", T, 1000, seed=11)).to(envA.device)
outs = envB.step_many(acts, per_step_outputs=True)
for t in range(T):
    o, r, d, f = envA.step(acts[t])
    do = (o - outs["obs"][t]).abs()
    if do.max() > 0 or (r - outs["reward"][t]).abs().max() > 0 or not torch.equal(d, outs["done"][t]):
        cols = (do > 0).any(0).nonzero().flatten().tolist()
        rows = (do > 0).any(1).nonzero().flatten().tolist()
        print("t", t, "max dobs", float(do.max()), "cols", cols, "nrows", len(rows), "drew", float((r - outs["reward"][t]).abs().max()),
              "done eq", torch.equal(d, outs["done"][t]))
        if t > 3: break
