"""The batched FP64 oracle (oracle/batched_oracle.py) is the same algorithm as the per-environment oracle, which is
pinned to the reference's own code: checked here step for step against both the per-environment oracle and the
reference-minted fixtures of the configurations it covers.  CPU only."""
import glob
import os

import numpy as np
import pytest

from oracle import dyn_oracle as O
from oracle.batched_oracle import BatchedOracle

HOVER = 0.092227


def _actions(mode, T, N, seed):
    u = np.random.default_rng(seed).uniform(-1, 1, size=(T, N, 4))
    a = {"saturating": u, "hover_band": HOVER + 0.002 * u, "mixed": HOVER + 0.006 * u, "physical": 0.06615 + 0.02 * u, "rpm": 0.6 * u}[mode]
    return a.astype(np.float32)


@pytest.mark.parametrize("track,S,mode,kw", [
    ("circle", 1, "saturating", {}), ("circle", 8, "mixed", {}), ("reaching", 8, "saturating", {}), ("reaching", 1, "hover_band", {}),
    ("circle", 8, "mixed", {"reward_id": "dummy"}), ("reaching", 8, "hover_band", {"reward_id": "thrustenv"}),
    ("circle", 8, "physical", {"normalize_actions": False, "include_distance": False}),
    ("circle", 8, "rpm", {"act": O.ACT_RPM, "drone_model": O.MODEL_CF2P, "normalize_actions": False}),
    ("reaching", 8, "rpm", {"act": O.ACT_ONE_D_RPM, "drone_model": O.MODEL_RACE, "normalize_actions": False}),
    ("circle", 1, "hover_band", {"max_steps": 25}),
    ("circle", 8, "mixed", {"normalize_obs": True}),
    ("reaching", 8, "saturating", {"normalize_obs": True, "normalize_reward": True, "clip_reward": 10.0}),
    ("circle", 1, "saturating", {"normalize_reward": True, "max_steps": 40}),
    # the documented DYN extensions: drag / ground effect (formulas pinned to the reference's own functions, test_ref_pins.py)
    # and the analytic ground-plane contact
    ("circle", 8, "mixed", {"physics": O.PHYSICS_DYN_GND_DRAG}), ("reaching", 8, "saturating", {"physics": O.PHYSICS_DYN_DRAG}),
    ("circle", 1, "hover_band", {"physics": O.PHYSICS_DYN_GND}),
    ("circle", 8, "saturating", {"ground_contact": True}), ("reaching", 8, "mixed", {"ground_contact": True, "physics": O.PHYSICS_DYN_GND_DRAG}),
])
def test_batched_oracle_equals_per_environment_oracle(track, S, mode, kw):
    N, T = 12, 240 if S == 1 else 80
    B = BatchedOracle(N, track, pyb_freq=240, ctrl_freq=240 // S, **kw)
    kw = dict(kw)
    wrap = dict(normalize_obs=kw.pop("normalize_obs", False), normalize_reward=kw.pop("normalize_reward", False),
                clip_reward=kw.pop("clip_reward", 0.0))
    ws = [O.OracleWorker(O.make_reference_env(track, pyb_freq=240, ctrl_freq=240 // S, **kw), **wrap) for _ in range(N)]
    obs0 = np.stack([w.reset()[0] for w in ws])
    np.testing.assert_allclose(B.reset_obs(), obs0, rtol=0, atol=1e-12 if wrap["normalize_obs"] else 0)
    otol = 1e-9 if wrap["normalize_obs"] else 2e-7
    acts = _actions(mode, T, N, seed=S + len(mode) + len(kw))
    dones = captures = 0
    for t in range(T):
        prev_idx = [w.env._current_target_index for w in ws]
        obs, rew, bits, found, term, ep_r, ep_l = B.step(acts[t])
        for i, w in enumerate(ws):
            o, r, d, info = w.step(acts[t, i])
            want = (1 if w.last_terminated else 0) | (2 if w.last_truncated else 0)
            assert bits[i] == want and found[i] == info["found_targets"], (t, i)
            np.testing.assert_allclose(obs[i], o, rtol=0, atol=otol)
            assert abs(rew[i] - float(r)) <= (1e-9 if wrap["normalize_reward"] else 1e-11) * max(1.0, abs(float(r))), (t, i, rew[i], r)
            if d:
                dones += 1
                np.testing.assert_allclose(term[i], info["terminal_observation"], rtol=0, atol=otol)
                assert ep_l[i] == info["episode"]["l"] and abs(ep_r[i] - info["episode"]["r"]) <= 1e-5 + 1e-9 * abs(ep_r[i])
            captures += int(info["found_targets"] > prev_idx[i])
            mm = min(B.margin[i], B.rew_margin[i])           # the per-environment oracle keeps one list for both kinds
            assert abs(mm - PU_min_margin(w.env)) <= 1e-9 * max(1.0, mm), (t, i, mm, PU_min_margin(w.env))
        st = B.state()
        for i, w in enumerate(ws):
            e = w.env
            np.testing.assert_allclose(st["pos"][i], e.pos, rtol=0, atol=1e-12)
            np.testing.assert_allclose(st["vel"][i], e.vel, rtol=0, atol=1e-11)
            np.testing.assert_allclose(st["rpy_rates"][i], e.rpy_rates, rtol=0, atol=1e-9)
            assert min(np.abs(st["quat"][i] - e.quat).max(), np.abs(st["quat"][i] + e.quat).max()) <= 1e-12
            assert abs(st["dist"][i] - e._distance_to_target) <= 1e-12 and abs(st["prev_dist"][i] - e._prev_distance_to_target) <= 1e-12
            assert st["target_idx"][i] == e._current_target_index and st["steps"][i] == e._steps and bool(st["just_found"][i]) == e.just_found
    assert dones > 0 or mode == "hover_band" or "ONE_D" in str(kw).upper()
    print(f"[{track} S={S} {mode} {kw}] dones={dones} captures={captures}")


def PU_min_margin(env):
    from tests.parity_utils import min_margin
    return min_margin(env)


_GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref_*.npz")))


def _covered(path):
    from tests.test_ref_golden import _meta
    m = _meta(np.load(path))
    return (m["reward"] in ("default", "dummy", "thrustenv") and not m["norm"] and not m["norm_rew"] and not m["clip_rew"]
            and m["act"] in ("thrust", "rpm", "one_d_rpm"))


_COVERED = [p for p in _GOLD if _covered(p)]


@pytest.mark.parametrize("path", _COVERED, ids=[os.path.basename(p)[:-4] for p in _COVERED])
def test_batched_oracle_matches_reference_fixtures(path):
    from tests.test_ref_golden import _meta
    g = np.load(path)
    m = _meta(g)
    T, N = g["reward"].shape
    B = BatchedOracle(N, m["track"], pyb_freq=240, ctrl_freq=240 // m["S"], max_steps=m["max_steps"], reward_id=m["reward"],
                      include_distance=m["include_distance"], normalize_actions=m["normalize_actions"], act=m["act"],
                      drone_model=m["model"], numpy_legacy_cast=False)
    np.testing.assert_allclose(B.reset_obs(), g["obs0"], rtol=0, atol=2e-7)
    for t in range(T):
        obs, rew, bits, found, term, ep_r, ep_l = B.step(g["actions"][t])
        np.testing.assert_array_equal(bits, g["done"][t])
        np.testing.assert_array_equal(found, g["found_targets"][t])
        np.testing.assert_allclose(obs, g["obs"][t], rtol=0, atol=2e-7)
        np.testing.assert_allclose(rew, g["reward"][t], rtol=1e-11, atol=1e-11)
        d = bits != 0
        np.testing.assert_allclose(term[d], g["terminal_obs"][t][d], rtol=0, atol=2e-7)
        np.testing.assert_array_equal(ep_l[d], g["ep_length"][t][d])
        np.testing.assert_allclose(ep_r[d], g["ep_return"][t][d], rtol=1e-9, atol=1e-9)
    assert len(_COVERED) >= 10
