"""Known-answer tests that pin the CPU oracle (oracle/dyn_oracle.py).

The reference holds no golden vectors for this path, so the pins are
(i) analytic fixed points / closed forms of the cited equations and (ii) the scenario
values listed in SURVEY.md section 8(c), re-derived here.
"""
import math

import numpy as np
import pytest

from oracle.dyn_oracle import (CF2X, OracleDroneEnv, OracleWorker, bullet_euler_from_quaternion,
                               bullet_matrix_from_quaternion, bullet_quaternion_from_euler,
                               make_reference_env, physical_action_bounds, rescale_action, thrust_to_rpm)

HOVER_ACTION = 0.0922265


def test_derived_constants():
    # BaseAviary.py:163-176 evaluated on safegym/cf2x.urdf
    assert CF2X.GRAVITY == pytest.approx(0.2646)
    assert CF2X.HOVER_RPM == pytest.approx(14468.429, abs=1e-3)
    assert CF2X.MAX_RPM == pytest.approx(21702.644, abs=1e-3)
    assert CF2X.GND_EFF_H_CLIP == pytest.approx(0.0377637, abs=1e-7)
    lo, hi = physical_action_bounds()
    assert lo.dtype == np.float32 and hi.dtype == np.float32
    assert lo[0] == np.float32(0.028161688) and hi[0] == np.float32(0.14834145)


def test_action_map_table():
    # PBDroneEnv.py:949-971,872-895 + env_utils.py:8-59 -- near bang-bang map
    b = physical_action_bounds()
    a = np.array([-1, 0.0935, 1, 0.08], np.float32)
    rpm = thrust_to_rpm(rescale_action(a, b), b)
    assert rpm.dtype == np.float32
    np.testing.assert_allclose(rpm, [9440.3, 16625.35, 21666.447, 9440.3], rtol=2e-7)
    # saturation edges of the band in which the policy action passes through unchanged
    for x, want in [(0.0899, 9440.3), (0.0972, 21666.447)]:
        r = thrust_to_rpm(rescale_action(np.full(4, x, np.float32), b), b)
        np.testing.assert_allclose(r, want, rtol=2e-7)


def test_batched_action_map_equals_literal():
    from oracle.dyn_oracle import rescale_action_batch, thrust_to_rpm_batch
    b = physical_action_bounds()
    rng = np.random.default_rng(3)
    a = np.concatenate([rng.uniform(-1, 1, (200, 4)), 0.0935 + 0.005 * rng.uniform(-1, 1, (400, 4))]).astype(np.float32)
    got = thrust_to_rpm_batch(rescale_action_batch(a, b), b)
    assert got.dtype == np.float32
    for i in range(a.shape[0]):
        ref = thrust_to_rpm(rescale_action(a[i], b), b)
        np.testing.assert_array_equal(got[i].view(np.int32), ref.view(np.int32))


def test_rotation_and_euler_roundtrip():
    rng = np.random.default_rng(0)
    for _ in range(200):
        rpy = rng.uniform([-3.1, -1.5, -3.1], [3.1, 1.5, 3.1])
        q = bullet_quaternion_from_euler(rpy)
        assert np.dot(q, q) == pytest.approx(1.0, abs=1e-14)
        np.testing.assert_allclose(bullet_euler_from_quaternion(q), rpy, atol=1e-9)
        R = bullet_matrix_from_quaternion(q)
        np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-13)
        # ZYX convention: forward vector (cos y cos p, sin y cos p, sin p) == (R00, R10, -R20)
        f = np.array([math.cos(rpy[2]) * math.cos(rpy[1]), math.sin(rpy[2]) * math.cos(rpy[1]), math.sin(rpy[1])])
        np.testing.assert_allclose([R[0, 0], R[1, 0], -R[2, 0]], f, atol=1e-12)


def test_euler_gimbal_branch():
    q = bullet_quaternion_from_euler([0.3, math.pi / 2, 0.1])
    r = bullet_euler_from_quaternion(q)
    assert r[0] == 0.0 and r[1] == pytest.approx(math.pi / 2)


def _free_env(**kw):
    return OracleDroneEnv(target_points=[[0, 0, 50.0]], threshold=0.3, discount=0.999, max_steps=10 ** 6,
                          aviary_dim=[-100, -100, 0, 100, 100, 100], initial_xyzs=[[0, 0, 1.0]],
                          cylinder=False, act="rpm", **kw)


def test_hover_fixed_point():
    # rpm == HOVER_RPM  =>  4 KF rpm^2 == M G: state constant (BaseAviary.py:922-925,165)
    env = _free_env()
    env.reset()
    for _ in range(240):
        env._dynamics(np.full(4, CF2X.HOVER_RPM))
    np.testing.assert_allclose(env.pos, [0, 0, 1], atol=1e-12)
    np.testing.assert_allclose(env.quat, [0, 0, 0, 1], atol=0)


def test_free_fall_semi_implicit_euler():
    # rpm = 0: v_n = -g n dt, z_n = z_0 - g dt^2 n(n+1)/2 (BaseAviary.py:941-943: pos uses the NEW vel)
    env = _free_env()
    env.reset()
    n, dt = 24, 1 / 240
    for _ in range(n):
        env._dynamics(np.zeros(4))
    assert env.pos[2] == pytest.approx(1 - 9.8 * dt * dt * n * (n + 1) / 2, abs=1e-13)
    assert env.pos[2] == pytest.approx(0.94895833333, abs=1e-10)
    assert env.vel[2] == pytest.approx(-9.8 * n * dt, abs=1e-13)


def test_pure_yaw_torque_sign_and_quaternion_norm():
    env = _free_env()
    env.reset()
    rpm = np.array([14000.0, 15000.0, 14000.0, 15000.0])   # props 1,3 faster -> +z torque (BaseAviary.py:929)
    for _ in range(120):
        env._dynamics(rpm)
    assert env.rpy_rates[2] > 0 and abs(env.rpy_rates[0]) < 1e-12 and abs(env.rpy_rates[1]) < 1e-12
    assert np.dot(env.quat, env.quat) == pytest.approx(1.0, abs=1e-15)
    rpm = np.array([15500.0, 14200.0, 13900.0, 14800.0])
    for _ in range(120):
        env._dynamics(rpm)
    assert np.dot(env.quat, env.quat) == pytest.approx(1.0, abs=1e-15)


def test_roll_torque_sign():
    # tau_x = (f0 + f1 - f2 - f3) L / sqrt2 (BaseAviary.py:931)
    env = _free_env()
    env.reset()
    env._dynamics(np.array([15000.0, 15000.0, 14000.0, 14000.0]))
    assert env.rpy_rates[0] > 0 and env.rpy_rates[1] == pytest.approx(0, abs=1e-15)


def test_reset_obs_and_hover_reward():
    env = make_reference_env("circle")
    obs, info = env.reset()
    assert obs.dtype == np.float32 and obs.shape == (13,)
    np.testing.assert_allclose(obs, [0.5, 0, 0.5] + [0] * 9 + [0.25], atol=0)
    a = np.full(4, HOVER_ACTION, np.float32)
    for _ in range(5):
        obs, r, term, trunc, info = env.step(a)
    assert env.pos[2] == pytest.approx(1.0, abs=1e-6)
    assert r == pytest.approx((3 * math.exp(-2) - 3) / 25, abs=1e-6)      # -0.1038: exp term + orientation penalty
    assert info["found_targets"] == 0 and not term and not trunc


def test_max_thrust_terminates_at_step_53_with_stale_reset():
    env = make_reference_env("circle")
    env.reset()
    a = np.ones(4, np.float32)
    for k in range(1, 200):
        obs, r, term, trunc, info = env.step(a)
        if term:
            break
    assert k == 53 and r == -10.0
    assert env.pos[2] == pytest.approx(1.3025, abs=1e-4)            # left the 0.3 m circle tube upwards
    obs, _ = env.reset()
    # reset obs is computed before the distances are reset, from the stale last post-step position
    assert obs[12] == pytest.approx(0.26039, abs=1e-5)
    assert env._distance_to_target == pytest.approx(1.04157, abs=1e-5)
    np.testing.assert_allclose(env._current_position, [1, 0, 1.29131], atol=1e-5)


def test_min_thrust_terminates_at_step_78():
    env = make_reference_env("circle")
    env.reset()
    a = -np.ones(4, np.float32)
    for k in range(1, 200):
        _, r, term, _, _ = env.step(a)
        if term:
            break
    assert k == 78 and r == -10.0 and env.pos[2] == pytest.approx(0.699, abs=1e-3)


def test_reaching_track_first_step_capture():
    env = make_reference_env("reaching")
    env.reset()
    a = np.full(4, HOVER_ACTION, np.float32)
    _, r, term, _, info = env.step(a)
    assert float(np.float32(r)) == pytest.approx(2.8, abs=1e-6) and info["found_targets"] == 1 and not term
    _, r, term, _, info = env.step(a)
    assert r == pytest.approx(-0.1193, abs=1e-4) and info["found_targets"] == 1
    assert env._distance_to_target == pytest.approx(2.53969, abs=1e-5)


def test_truncation_on_step_max_plus_one():
    env = make_reference_env("circle", max_steps=7)
    env.reset()
    a = np.full(4, HOVER_ACTION, np.float32)
    flags = [env.step(a)[3] for _ in range(8)]
    assert flags == [False] * 7 + [True]          # _steps is read before the increment (PBDroneEnv.py:444-454)


def test_final_target_reward_and_done():
    targets = [[1.0, 0.0, 1.0]]
    env = OracleDroneEnv(target_points=targets, threshold=0.3, discount=0.999, max_steps=100,
                         aviary_dim=[-2, -2, 0, 2, 2, 2], initial_xyzs=[[1, 0, 1]], circle=True,
                         include_distance=True, normalize_actions=True)
    env.reset()
    _, r, term, _, info = env.step(np.full(4, HOVER_ACTION, np.float32))
    assert float(r) == 8.0 and term and info["found_targets"] == 1


def test_worker_autoreset_contract():
    w = OracleWorker(make_reference_env("circle"), normalize_obs=False)
    w.reset()
    a = np.ones(4, np.float32)
    for k in range(1, 60):
        obs, r, done, info = w.step(a)
        if done:
            break
    assert k == 53 and info["episode"]["l"] == 53 and not info["TimeLimit.truncated"]
    assert info["terminal_observation"].shape == (13,)
    np.testing.assert_allclose(obs[:3], [0.5, 0, 0.5])          # returned obs is the reset obs
    assert obs[12] == pytest.approx(0.26039, abs=1e-5)
    assert info["episode"]["r"] == pytest.approx(-10 + 52 * np.mean([0]) + sum([0]), abs=20)  # finite, crash-dominated


def test_normalize_observation_running_stats():
    # normalize.py:10-47 with batch size 1: after k identical samples x the mean is x k/(k+1e-4)
    w = OracleWorker(make_reference_env("circle"), normalize_obs=True)
    o = w.reset()[0]
    assert w.obs_rms.count == pytest.approx(1.0001)
    assert w.obs_rms.mean[0] == pytest.approx(0.5 / 1.0001)
    assert o.shape == (13,)
