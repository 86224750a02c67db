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


# ---- airframes, controller and literature rewards: analytic pins of the restatements added for SURVEY f3 / a19 -----------
def test_airframe_constants_and_torque_mix_signs():
    """Sol/resources/cf2p.urdf / racer.urdf constants and the torque-mix branches of BaseAviary._dynamics (:927-935)."""
    from oracle.dyn_oracle import CF2P, RACE, ACT_RPM, MODEL_CF2P, MODEL_RACE
    assert (CF2P.IXX, CF2P.IZZ, CF2P.L, CF2P.M) == (2.3951e-5, 3.2347e-5, 0.0397, 0.027)
    assert (RACE.M, RACE.L, RACE.KF, RACE.KM, RACE.THRUST2WEIGHT_RATIO) == (0.830, 0.109, 8.47e-9, 2.13e-11, 4.17)
    assert RACE.HOVER_RPM == pytest.approx(math.sqrt(9.8 * 0.830 / (4 * 8.47e-9)))
    # one substep from rest with a single motor sped up: sign and size of the body-rate change per airframe
    for model, c in ((MODEL_CF2P, CF2P), (MODEL_RACE, RACE)):
        for motor in range(4):
            e = make_reference_env("circle", act=ACT_RPM, drone_model=model, normalize_actions=False)
            e.reset()
            rpm = np.full(4, c.HOVER_RPM)
            rpm[motor] *= 1.1
            e._dynamics(rpm)
            df = c.KF * (rpm[motor] ** 2 - c.HOVER_RPM ** 2)
            dz = c.KM * (rpm[motor] ** 2 - c.HOVER_RPM ** 2) * (1 if motor in (1, 3) else -1) * (-1 if model == MODEL_RACE else 1)
            if model == MODEL_CF2P:       # + frame: motor 1 / 3 on the +-y arm roll, motor 0 / 2 on the +-x arm pitch (:933-935)
                tx = {1: df, 3: -df}.get(motor, 0.0) * c.L
                ty = {0: -df, 2: df}.get(motor, 0.0) * c.L
            else:                         # x frame (:930-932)
                tx = (df if motor in (0, 1) else -df) * c.L / math.sqrt(2)
                ty = (df if motor in (1, 2) else -df) * c.L / math.sqrt(2)
            want = np.array([tx / c.IXX, ty / c.IYY, dz / c.IZZ]) / 240
            np.testing.assert_allclose(e.rpy_rates, want, rtol=1e-9, atol=1e-12)


def test_dslpid_hover_fixed_point_and_mixer():
    """DSLPIDControl at rest on its target with zero integrals commands exactly the hover RPM on all four motors
    (thrust = M g along body z; DSLPIDControl.py:164-195,225-261); a pure yaw error moves the motor pairs as the CF2X
    mixer's third column says (:47-53)."""
    from oracle.dyn_oracle import OracleDSLPID
    c = OracleDSLPID()
    q0 = np.array([0.0, 0.0, 0.0, 1.0])
    rpm, pos_e, yaw_e = c.computeControl(1 / 30, np.array([1.0, 0, 1]), q0, np.zeros(3), np.zeros(3), np.array([1.0, 0, 1]))
    np.testing.assert_allclose(rpm, CF2X.HOVER_RPM, rtol=1e-12)
    assert np.all(pos_e == 0) and yaw_e == 0
    assert np.all(c.integral_pos_e == 0) and np.all(c.integral_rpy_e == 0)
    # 0.1 rad of yaw: rot_e = (0, 0, 2 sin(0.1)); torque_z = -60000 rot_e_z + 12000 * (-(0.1 - 0) / ct) + 500 * integral
    c = OracleDSLPID()
    qy = np.array([0.0, 0.0, math.sin(0.05), math.cos(0.05)])
    rpm, _, yaw_e = c.computeControl(1 / 30, np.zeros(3), qy, np.zeros(3), np.zeros(3), np.zeros(3))
    assert yaw_e == pytest.approx(-0.1)
    e2 = 2 * math.sin(0.1)
    tz = max(-3200.0, -60000 * e2 + 12000 * (-0.1 * 30) + 500 * (-e2 / 30))
    base = (math.sqrt(CF2X.GRAVITY / (4 * CF2X.KF)) - 4070.3) / 0.2685
    want = 0.2685 * np.clip(base + np.array([-1, 1, -1, 1]) * tz, 20000, 65535) + 4070.3
    np.testing.assert_allclose(rpm, want, rtol=1e-12)
    np.testing.assert_allclose(c.last_rpy, [0, 0, 0.1], atol=1e-15)


def test_one_d_pid_holds_altitude_and_pid_family_never_resets_the_controller():
    from oracle.dyn_oracle import ACT_ONE_D_PID
    e = make_reference_env("circle", pyb_freq=240, ctrl_freq=30, act=ACT_ONE_D_PID, normalize_actions=False)
    e.reset()
    for _ in range(120):
        e.step(np.zeros(4, np.float32))           # target = current position: hover in place
    assert abs(e.pos[2] - 1.0) < 1e-9 and np.abs(e.vel).max() < 1e-9
    e.step(np.array([1.0, 0, 0, 0], np.float32))  # 0.1 m up: integral_pos_e[2] = 0.1 / 30
    assert e.ctrl.integral_pos_e[2] == pytest.approx(0.1 / 30)
    kept = e.ctrl.integral_pos_e.copy()
    e.reset()
    np.testing.assert_array_equal(e.ctrl.integral_pos_e, kept)      # BaseSingleAgentAviary never calls ctrl.reset()


def test_literature_reward_first_step_values():
    """Rewarder.py:66-150 by hand for the first step from the spawn point of the circle track: prev_dis = dis = 1 (stale pair),
    a_t-1 = 0, no capture, no crash; delta_cam = angle between (1, 0, 0) and the direction to target 0 = (0.5, 0.866, 1) - pos."""
    a = np.full(4, HOVER_ACTION, np.float32)
    for rid in ("bootstrapped", "champ"):
        e = make_reference_env("circle", reward_id=rid)
        e.reset()
        _, r, term, _, _ = e.step(a)
        v = e._target_points[0] - e.pos
        dc = math.acos(np.clip(np.dot(e.get_forward_vector(), v / np.linalg.norm(v)), -1, 1))
        da, w = np.linalg.norm(np.float64(a)), np.linalg.norm(e.rpy_rates)
        if rid == "bootstrapped":
            want = 0.5 * 0.0 + 0.025 * (2e-4 * dc ** 4) - 2e-4 * da - 5e-4 * w
        else:
            want = 1.0 * 0.0 + 0.02 * math.exp(-10 * dc ** 4) - 2e-4 * w ** 2 - 1e-4 * da ** 2
        assert not term and float(r) == pytest.approx(want, rel=1e-12, abs=1e-15)
        assert dc == pytest.approx(math.acos(-0.5 / 1.0), abs=2e-3)     # target 0 of the circle sits 120 degrees off the nose
