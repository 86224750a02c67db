"""Golden fixtures (tests/golden/*.npz, minted by tests/golden/make_golden.py from the oracle):
 - not gpu: the oracle still reproduces them bit for bit, and the host build of the device logic matches them;
 - gpu:     the CUDA path, through the C ABI, matches them within the stated tolerances."""
import glob
import os

import numpy as np
import pytest

from tests import parity_utils as PU
from tests.golden.make_golden import CASES, run

GOLD = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
              if not os.path.basename(p).startswith(("ref_", "forces_")))      # ref_*: tests/test_ref_golden.py, forces_ref: tests/test_ref_pins.py
PHYS = {"dyn": 0, "dyn_drag": 1, "dyn_gnd": 2, "dyn_gnd_drag": 3}


def test_fixtures_exist():
    assert len(GOLD) == len(CASES)


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_reproduces_golden(path):
    g = np.load(path)
    out = run(*CASES[os.path.basename(path)[:-4]])
    for k in ("obs0", "obs", "reward", "done", "found_targets"):
        np.testing.assert_array_equal(out[k], g[k])


def _compare(g, step_fn, obs0):
    np.testing.assert_allclose(obs0, g["obs0"], atol=1e-6)
    T, N = g["reward"].shape
    worst_obs = worst_rew = 0.0
    for t in range(T):
        o, r, d, f = step_fn(g["actions"][t])
        np.testing.assert_array_equal(d, g["done"][t])
        np.testing.assert_array_equal(f, g["found_targets"][t])
        for i in range(N):
            e = np.abs(o[i].astype(np.float64) - g["obs"][t, i])
            e[3:6] = np.minimum(e[3:6], np.abs(2 - e[3:6]))
            worst_obs = max(worst_obs, e[:9].max(), e[12])
        worst_rew = max(worst_rew, float(np.abs(r - g["reward"][t]).max()))
    # open-loop over the whole fixture (up to 480 substeps, no re-synchronisation): 10x the 1 s horizon tolerance
    assert worst_obs < 1e-3 and worst_rew < 1e-2, (worst_obs, worst_rew)
    return worst_obs, worst_rew


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_device_logic_matches_golden(path):
    from oracle.dyn_oracle import make_reference_env
    from tests.host_emu import HostEmuEnv
    g = np.load(path)
    track, S, mode, physics = [str(x) for x in g["meta"]]
    ref = make_reference_env(track)
    env = HostEmuEnv(g["reward"].shape[1], ref._target_points, aviary_dim=ref._aviary_dim, initial_xyzs=ref.INIT_XYZS,
                     pyb_freq=240, ctrl_freq=240 // int(S), circle=(track == "circle"), include_distance=True,
                     normalize_actions=True, physics=PHYS[physics])
    print(_compare(g, env.step, g["obs0"]))


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_cuda_matches_golden(path):
    import torch
    from drl_dronenavigation_b200 import Physics
    from drl_dronenavigation_b200.batched_env import BatchedDroneEnv
    from oracle.dyn_oracle import make_reference_env
    g = np.load(path)
    track, S, mode, physics = [str(x) for x in g["meta"]]
    ref = make_reference_env(track)
    phys = {"dyn": Physics.DYN, "dyn_drag": Physics.PYB_DRAG, "dyn_gnd": Physics.PYB_GND, "dyn_gnd_drag": Physics.PYB_GND_DRAG_DW}[physics]
    env = BatchedDroneEnv(g["reward"].shape[1], ref._target_points, aviary_dim=ref._aviary_dim, initial_xyzs=ref.INIT_XYZS,
                          pyb_freq=240, ctrl_freq=240 // int(S), circle=(track == "circle"), include_distance=True,
                          normalize_actions=True, physics=phys)
    obs0 = env.reset().cpu().numpy()

    def step(a):
        o, r, d, f = env.step(torch.from_numpy(a).cuda())
        return o.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy(), f.cpu().numpy()
    print(_compare(g, step, obs0))
    env.close()
