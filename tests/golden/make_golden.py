"""Mints the golden trajectories under tests/golden/ from the CPU oracle (oracle/dyn_oracle.py).

The reference holds no golden vectors for this path and cannot be imported in
this container (no pybullet / gymnasium / SB3, no network), so the fixtures are outputs of the
oracle restatement, which is itself pinned by the analytic tests of tests/test_oracle_kat.py.
Run:  python tests/golden/make_golden.py      (rewrites the .npz files deterministically)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.dyn_oracle import OracleWorker, make_reference_env  # noqa: E402

HOVER = 0.092227
CASES = {
    # name: (track, S, action mode, envs, steps, seed, physics)
    "circle_s1_saturating": ("circle", 1, "saturating", 4, 160, 11, "dyn"),
    "circle_s8_mixed": ("circle", 8, "mixed", 4, 60, 12, "dyn"),
    "reaching_s8_saturating": ("reaching", 8, "saturating", 4, 50, 13, "dyn"),
    "reaching_s1_hover": ("reaching", 1, "hover_band", 2, 200, 14, "dyn"),
    "circle_s8_drag_gnd": ("circle", 8, "mixed", 2, 40, 15, "dyn_gnd_drag"),
}


def actions(mode, T, N, seed):
    u = np.random.default_rng(seed).uniform(-1, 1, size=(T, N, 4))
    return {"saturating": u, "hover_band": HOVER + 0.002 * u, "mixed": HOVER + 0.006 * u}[mode].astype(np.float32)


def run(track, S, mode, N, T, seed, physics):
    ws = [OracleWorker(make_reference_env(track, pyb_freq=240, ctrl_freq=240 // S, physics=physics), normalize_obs=False)
          for _ in range(N)]
    a = actions(mode, T, N, seed)
    obs0 = np.stack([w.reset()[0] for w in ws])
    obs, rew, done, found = (np.zeros((T, N, 13), np.float32), np.zeros((T, N), np.float32),
                             np.zeros((T, N), np.uint8), np.zeros((T, N), np.int32))
    pos = np.zeros((T, N, 3))
    for t in range(T):
        for i, w in enumerate(ws):
            o, r, d, info = w.step(a[t, i])
            obs[t, i], rew[t, i], found[t, i] = o, np.float32(r), info["found_targets"]
            done[t, i] = (1 if w.last_terminated else 0) | (2 if w.last_truncated else 0)
            pos[t, i] = w.env.pos
    return dict(actions=a, obs0=obs0, obs=obs, reward=rew, done=done, found_targets=found, pos=pos,
                meta=np.array([track, str(S), mode, physics]))


if __name__ == "__main__":
    for name, cfg in CASES.items():
        out = run(*cfg)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "dones", int((out["done"] != 0).sum()), "captures", int(out["found_targets"].max()))
