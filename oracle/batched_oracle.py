"""CPU ORACLE, BATCHED FORM (test infrastructure, NOT product code).

The same restatement as ``oracle/dyn_oracle.py`` -- ``PBDroneEnv.step`` over ``BaseAviary._dynamics`` with the
``Monitor`` / ``SubprocVecEnv`` auto-reset contract around it -- written as numpy array operations over N
environments at once, so that (i) the CUDA path can be checked against an FP64 oracle at BASELINE.json's FULL sizes
(4096 environments for hundreds of control steps take seconds), and (ii) ``bench.py`` can quote a best-case CPU
figure next to the per-environment one (the reference itself steps one Python env object per worker process).

Only ``tests/`` and ``bench.py``'s ``cpu_baseline`` leg may import this file.

PARITY PIN: ``tests/test_batched_oracle.py`` checks this file (a) against the per-environment oracle, which is pinned
to the reference's own code (see the header of ``dyn_oracle.py``), step for step on shared seeded actions -- done bits,
found_targets and episode lengths exact, observations identical as float32 up to 1 ulp, rewards and state to 1e-12 --
and (b) directly against the reference-minted fixtures ``tests/golden/ref_*.npz`` of the configurations it covers.

Coverage (what the full-size parity tests need): CF2X / CF2P / RACE; THRUST (+ ``normalize_actions``) and RPM /
ONE_D_RPM action maps; ``Physics.DYN``; 12 / 13-dim observation; the PBDroneEnv reward family (``default``, ``dummy``,
``thrustenv``); circle and segment-tube termination; truncation; deterministic reset with the stale-distance quirks.
Not covered here (per-environment oracle only): drag / ground effect, the other reward families, random spawns, the
PID action types.  The wrappers of ``make_env`` (NormalizeObservation, reward clip, NormalizeReward; ``normalize.py:10-147``)
are included, per environment like the reference's worker processes.

Line citations (all under ``/root/reference``) are the ones of ``dyn_oracle.py``; each method names its counterpart.
"""
from __future__ import annotations

import numpy as np

from oracle import dyn_oracle as O

_REWARDS = {   # crash, final, capture, capture_orient, progress_w, orient_w, smoothness thresholds (dyn_oracle._computeReward)
    "default": (-10.0, 200.0, 75.0, 5.0, 3000.0, 3.0, (0.7, 0.3)),
    "dummy": (-10.0, 200.0, 75.0, 5.0, 3000.0, 3.0, (0.1, 0.1)),
    "thrustenv": (-4.0, 1000.0, 25.0, 0.0, 20.0, 0.0, None),
}


def _norm(v):
    return np.sqrt(np.sum(v * v, axis=-1))


class BatchedOracle:
    """N independent reference environments, state as [N, ...] float64 arrays named like the reference's attributes."""

    def __init__(self, num_envs, track="circle", pyb_freq=240, ctrl_freq=240, max_steps=4096, act=O.ACT_THRUST,
                 normalize_actions=True, include_distance=True, reward_id="default", threshold=0.3, cylinder=True,
                 drone_model=O.MODEL_CF2X, numpy_legacy_cast=True, normalize_obs=False, normalize_reward=False, clip_reward=0.0,
                 reward_gamma=0.99, physics=O.PHYSICS_DYN, ground_contact=False):
        if reward_id not in _REWARDS:
            raise ValueError(f"reward {reward_id!r} is only in the per-environment oracle")
        if act not in (O.ACT_THRUST, O.ACT_RPM, O.ACT_ONE_D_RPM):
            raise ValueError(f"action type {act!r} is only in the per-environment oracle")
        self.N = int(num_envs)
        # the documented DYN extensions: drag and ground effect (formulas pinned against the reference's own
        # functions by tests/test_ref_pins.py) and the analytic ground-plane contact
        self.drag = physics in (O.PHYSICS_DYN_DRAG, O.PHYSICS_DYN_GND_DRAG)
        self.gnd = physics in (O.PHYSICS_DYN_GND, O.PHYSICS_DYN_GND_DRAG)
        self.ground_contact = bool(ground_contact)
        self.C = O.AIRFRAMES[drone_model]
        self.DRONE_MODEL = drone_model
        self.targets, init, dim = O.circle_track() if track == "circle" else O.reaching_track()
        self.circle = (track == "circle")
        self.cylinder = cylinder
        self.T = len(self.targets)
        self.INIT_XYZ = np.array(init, dtype=np.float64).reshape(3)
        self.x_low, self.y_low, self.z_low, self.x_high, self.y_high, self.z_high = [float(v) for v in dim]
        self.max_target_dist = max(abs(self.x_low) + self.x_high, abs(self.y_low) + self.y_high, self.z_high)
        self.S = pyb_freq // ctrl_freq
        self.dt = 1.0 / pyb_freq
        self.max_steps, self.threshold = int(max_steps), float(threshold)
        self.act, self.normalize_actions, self.include_distance = act, normalize_actions, include_distance
        self.numpy_legacy_cast = numpy_legacy_cast
        self.rw = _REWARDS[reward_id]
        self.bounds = O.physical_action_bounds(self.C)
        self.obs_dim = 13 if include_distance else 12
        N = self.N
        # segment table of the non-circle tube (PBDroneEnv.py:746-786): base1 -> base2 per target index
        b1 = np.vstack([self.INIT_XYZ[None], self.targets[:-1]])
        self._seg_b1, self._seg_b2 = b1, self.targets.copy()
        # BaseAviary._housekeeping + PBDroneEnv.__init__ tail
        self.q0 = O.bullet_pose_readback(O.bullet_quaternion_from_euler(np.zeros(3)))
        self.pos = np.tile(self.INIT_XYZ, (N, 1))
        self.quat = np.tile(self.q0, (N, 1))
        self.rpy = np.tile(O.bullet_euler_from_quaternion(self.q0), (N, 1))
        self.vel, self.ang_v, self.rpy_rates = np.zeros((N, 3)), np.zeros((N, 3)), np.zeros((N, 3))
        self.cur_pos = self.pos.copy()                                   # PBDroneEnv._current_position
        self.current_vel, self.current_ang_v = np.zeros((N, 3)), np.zeros((N, 3))
        self.prev_vel, self.prev_ang_v = np.zeros((N, 3)), np.zeros((N, 3))
        d0 = _norm(self.cur_pos - self.targets[0])
        self.dist, self.prev_dist = d0.copy(), d0.copy()
        self.idx = np.zeros(N, np.int64)
        self.steps = np.zeros(N, np.int64)
        self.just_found = np.zeros(N, bool)
        self.ep_return, self.ep_len = np.zeros(N), np.zeros(N, np.int64)
        self.last_rpm = np.zeros((N, 4))
        self.last_clipped = np.zeros((N, 4))    # BaseAviary.last_clipped_action (drag reads the PREVIOUS substep's), float64 zeros at reset
        # the wrappers of PBDroneSimulator.make_env (:181-195), per environment as each worker process has its own:
        # NormalizeObservation / NormalizeReward (normalize.py:50-147) with RunningMeanStd batches of one, TransformReward clip
        self.normalize_obs, self.normalize_reward, self.clip_reward, self.reward_gamma = normalize_obs, normalize_reward, clip_reward, reward_gamma
        D = self.obs_dim
        self.obs_mean, self.obs_var, self.obs_count = np.zeros((N, D)), np.ones((N, D)), np.full(N, 1e-4)
        self.ret_mean, self.ret_var, self.ret_count, self.returns = np.zeros(N), np.ones(N), np.full(N, 1e-4), np.zeros(N)
        self.margin = np.full(N, np.inf)       # test instrumentation: distance of this step's discrete decisions to their thresholds
        self.rew_margin = np.full(N, np.inf)   # the same for decisions that only change the reward (orientation, smoothness)
        self.gimbal_margin = np.full(N, np.inf)

    # ---- helpers ------------------------------------------------------------------------------------------------
    @staticmethod
    def _rms_update(x, mean, var, count, sel):
        """normalize.RunningMeanStd.update with a batch of one (normalize.py:19-47) for the rows `sel`; in place."""
        lead = (slice(None),) + (None,) * (x.ndim - 1)
        cnt = count[lead]
        batch_mean, batch_var, batch_count = x, 0.0, 1             # np.mean / np.var of a batch of one
        delta = batch_mean - mean
        tot = cnt + batch_count
        new_mean = mean + delta * batch_count / tot
        M2 = var * cnt + batch_var * batch_count + np.square(delta) * cnt * batch_count / tot
        new_var = M2 / tot
        s_ = sel[lead]
        mean[...] = np.where(s_, new_mean, mean)
        var[...] = np.where(s_, new_var, var)
        tot = count + batch_count
        count[...] = np.where(sel, tot, count)

    def _normalize_obs(self, obs, sel):
        """NormalizeObservation.normalize (normalize.py:94-97): update, then (obs - mean) / sqrt(var + 1e-8); rows `sel`."""
        x = obs.astype(np.float64)
        self._rms_update(x, self.obs_mean, self.obs_var, self.obs_count, sel)
        return np.where(sel[:, None], (x - self.obs_mean) / np.sqrt(self.obs_var + 1e-8), x)

    def _m(self, x, into="margin"):
        with np.errstate(invalid="ignore"):
            a = np.abs(x)
        a = np.where(np.isfinite(a), a, np.inf)
        setattr(self, into, np.minimum(getattr(self, into), a))

    @staticmethod
    def _rotation(q):
        """p.getMatrixFromQuaternion for [N, 4] (btMatrix3x3::setRotation)."""
        x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
        s = 2.0 / (x * x + y * y + z * z + w * w)
        xs, ys, zs = x * s, y * s, z * s
        wx, wy, wz = w * xs, w * ys, w * zs
        xx, xy, xz = x * xs, x * ys, x * zs
        yy, yz, zz = y * ys, y * zs, z * zs
        R = np.empty((q.shape[0], 3, 3))
        R[:, 0, 0], R[:, 0, 1], R[:, 0, 2] = 1.0 - (yy + zz), xy - wz, xz + wy
        R[:, 1, 0], R[:, 1, 1], R[:, 1, 2] = xy + wz, 1.0 - (xx + zz), yz - wx
        R[:, 2, 0], R[:, 2, 1], R[:, 2, 2] = xz - wy, yz + wx, 1.0 - (xx + yy)
        return R

    @staticmethod
    def _euler(q):
        """p.getEulerFromQuaternion for [N, 4] (pybullet.c, gimbal branches at |sarg| >= 0.99999)."""
        x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
        sqx, sqy, sqz, squ = x * x, y * y, z * z, w * w
        sarg = -2.0 * (x * z - w * y)
        lo, hi = sarg <= -0.99999, sarg >= 0.99999
        roll = np.arctan2(2.0 * (y * z + w * x), squ - sqx - sqy + sqz)
        pitch = np.arcsin(np.clip(sarg, -1.0, 1.0))
        yaw = np.arctan2(2.0 * (x * y + w * z), squ + sqx - sqy - sqz)
        roll = np.where(lo | hi, 0.0, roll)
        pitch = np.where(lo, -0.5 * np.pi, np.where(hi, 0.5 * np.pi, pitch))
        yaw = np.where(lo, 2.0 * np.arctan2(x, -y), np.where(hi, 2.0 * np.arctan2(-x, y), yaw))
        return np.stack([roll, pitch, yaw], axis=1)

    # ---- action -> rpm (dyn_oracle.rescale_action / thrust_to_rpm / rpm_action_to_rpm) -------------------------------
    def _preprocess(self, actions):
        a = np.asarray(actions, dtype=np.float32).reshape(self.N, 4)
        if self.act == O.ACT_THRUST:
            if self.normalize_actions:
                a = O.rescale_action_batch(a, self.bounds)
            return O.thrust_to_rpm_batch(a, self.bounds, self.C)
        rpm = O.rpm_action_to_rpm(a, self.C, self.numpy_legacy_cast)
        if self.act == O.ACT_ONE_D_RPM:
            rpm = np.repeat(rpm[:, :1], 4, axis=1)
        return rpm

    # ---- BaseAviary._dynamics + _integrateQ (dyn_oracle._dynamics) ----------------------------------------------------
    def _dynamics(self, rpm):
        c, dt = self.C, self.dt
        R = self._rotation(self.quat)
        forces = np.array(rpm ** 2) * c.KF                          # float32 on the THRUST path
        if self.gnd:                                                # dyn_oracle._ground_effect (BaseAviary.py:798-834)
            rpy = self._euler(self.quat)
            offs = np.array(c.PROP_XY, dtype=np.float64)            # [4, 2]
            heights = self.pos[:, 2:3] + R[:, 2, 0:1] * offs[None, :, 0] + R[:, 2, 1:2] * offs[None, :, 1]
            heights = np.clip(heights, c.GND_EFF_H_CLIP, np.inf)
            gnd = np.array(rpm, dtype=np.float64) ** 2 * c.KF * c.GND_EFF_COEFF * (c.PROP_RADIUS / (4 * heights)) ** 2
            ok = (np.abs(rpy[:, 0]) < np.pi / 2) & (np.abs(rpy[:, 1]) < np.pi / 2)
            forces = forces + np.where(ok[:, None], gnd, 0.0)
        total = forces[:, 0] + forces[:, 1] + forces[:, 2] + forces[:, 3]      # np.sum over four elements, left to right
        thrust_world = R[:, :, 2] * np.float64(total)[:, None]
        force_world = thrust_world - np.array([0, 0, c.GRAVITY])
        if self.drag:                                               # dyn_oracle._drag (BaseAviary.py:838-865), LINK_FRAME: rotated twice
            last = self.last_clipped
            per = 2 * np.pi * last / 60                             # float32 arithmetic when `last` is the float32 action-map output
            ssum = per[:, 0] + per[:, 1] + per[:, 2] + per[:, 3]
            factors = -1 * np.array([c.DRAG_COEFF_XY, c.DRAG_COEFF_XY, c.DRAG_COEFF_Z])[None, :] * np.float64(ssum)[:, None]
            link = np.einsum("nij,nj->ni", R, factors * self.vel)
            force_world = force_world + np.einsum("nij,nj->ni", R, link)
        z_t = np.array(rpm ** 2) * c.KM
        if self.DRONE_MODEL == O.MODEL_RACE:
            z_t = -z_t
        z_torque = (-z_t[:, 0] + z_t[:, 1] - z_t[:, 2] + z_t[:, 3])
        if self.DRONE_MODEL in (O.MODEL_CF2X, O.MODEL_RACE):
            x_torque = (forces[:, 0] + forces[:, 1] - forces[:, 2] - forces[:, 3]) * (c.L / np.sqrt(2))
            y_torque = (-forces[:, 0] + forces[:, 1] + forces[:, 2] - forces[:, 3]) * (c.L / np.sqrt(2))
        else:
            x_torque = (forces[:, 1] - forces[:, 3]) * c.L
            y_torque = (-forces[:, 0] + forces[:, 2]) * c.L
        torques = np.stack([np.float64(x_torque), np.float64(y_torque), np.float64(z_torque)], axis=1)
        w = self.rpy_rates
        Jw = w * np.array([c.IXX, c.IYY, c.IZZ])
        torques = torques - np.cross(w, Jw)
        deriv = torques * np.diag(c.J_INV)
        self.vel = self.vel + dt * (force_world / c.M)
        w = w + dt * deriv
        self.pos = self.pos + dt * self.vel
        # _integrateQ (BaseAviary.py:960-973)
        n = _norm(w)
        still = np.isclose(n, 0)
        p_, q_, r_ = w[:, 0], w[:, 1], w[:, 2]
        Q = self.quat
        lam = 0.5 * np.stack([r_ * Q[:, 1] - q_ * Q[:, 2] + p_ * Q[:, 3],
                              -r_ * Q[:, 0] + p_ * Q[:, 2] + q_ * Q[:, 3],
                              q_ * Q[:, 0] - p_ * Q[:, 1] + r_ * Q[:, 3],
                              -p_ * Q[:, 0] - q_ * Q[:, 1] - r_ * Q[:, 2]], axis=1)
        theta = n * dt / 2
        with np.errstate(all="ignore"):
            newq = Q * np.cos(theta)[:, None] + (2 / n * np.sin(theta))[:, None] * lam
        newq = np.where(still[:, None], Q, newq)
        self.quat = newq / _norm(newq)[:, None]                     # Bullet pose read-back
        self.ang_v = np.einsum("nij,nj->ni", R, w)
        self.rpy_rates = w
        self.last_clipped = np.array(rpm)                            # BaseAviary.py:442, inside the substep loop

    # ---- PBDroneEnv._computeObs (dyn_oracle._computeObs) ---------------------------------------------------------------
    def _obs(self, pos, rpy, vel, ang_v, dist):
        n = _norm(ang_v)
        with np.errstate(all="ignore"):
            ang = np.where((n != 0)[:, None], ang_v / n[:, None], ang_v)
        cols = [pos[:, 0] / self.x_high, pos[:, 1] / self.y_high, pos[:, 2] / self.z_high,
                np.clip(rpy[:, 0], -np.pi, np.pi) / np.pi, np.clip(rpy[:, 1], -np.pi, np.pi) / np.pi, rpy[:, 2] / np.pi,
                np.clip(vel[:, 0], -3, 3) / 3, np.clip(vel[:, 1], -3, 3) / 3, np.clip(vel[:, 2], -1, 1) / 3,
                ang[:, 0], ang[:, 1], ang[:, 2]]
        if self.include_distance:
            cols.append(dist / self.max_target_dist)
        ret = np.stack(cols, axis=1)
        return np.clip(ret, np.finfo(np.float32).min, np.finfo(np.float32).max).astype(np.float32)

    # ---- PBDroneEnv._has_collision_occurred / is_out_of_cylinder_bounds -------------------------------------------------
    def _collided(self, idx, record=True):
        p = self.pos
        if record:
            for m in (p[:, 0] - self.x_high, p[:, 0] - self.x_low, p[:, 1] - self.y_high, p[:, 1] - self.y_low, p[:, 2] - self.z_high):
                self._m(m)
        out = (p[:, 0] > self.x_high) | (p[:, 0] < self.x_low) | (p[:, 1] > self.y_high) | (p[:, 1] < self.y_low) | (p[:, 2] > self.z_high)
        if self.ground_contact:                                     # analytic substitute for p.getContactPoints() (dyn_oracle._has_collision_occurred)
            if record:
                self._m(p[:, 2] - self.C.COLLISION_H / 2)
            out = out | (p[:, 2] < self.C.COLLISION_H / 2)
        if not self.cylinder:
            return out
        if self.circle:
            c2d = p - np.array([0.0, 0.0, 1.0])
            c2d[:, 2] = 0
            with np.errstate(all="ignore"):
                closest = np.array([0.0, 0.0, 1.0]) + c2d / _norm(c2d)[:, None] * 1
                d = _norm(p - closest)
            if record:
                self._m(d - self.threshold)
            with np.errstate(invalid="ignore"):
                return out | (d > self.threshold)
        k = np.minimum(idx, self.T - 1)
        b1, b2 = self._seg_b1[k], self._seg_b2[k]
        line = b2 - b1
        ll = _norm(line)
        zero = (ll == 0)
        with np.errstate(all="ignore"):
            u = line / ll[:, None]
        u = np.where(zero[:, None], 0.0, u)
        e1, e2 = b1 - 0.2 * u, b2 + 0.2 * u
        proj = np.clip(np.sum((p - e1) * u, axis=1), 0, _norm(e2 - e1))
        d_line = _norm(p - (e1 + proj[:, None] * u))
        d_zero = _norm(p - b1)
        d = np.where(zero, d_zero - self.threshold, d_line - (self.threshold + 0.2))
        if record:
            self._m(d)
        return out | (d > 0)

    # ---- one control step of every environment ---------------------------------------------------------------------------
    def step(self, actions):
        """Returns (obs [N, D] f32 -- the reset obs where done --, reward [N] f64, done bits [N] u8, found_targets [N],
        terminal_obs [N, D] f32 (rows valid where done), episode_return [N], episode_length [N])."""
        N, T = self.N, self.T
        self.margin[:] = np.inf
        self.rew_margin[:] = np.inf
        rpm = self._preprocess(actions)
        for _ in range(self.S):
            self._dynamics(rpm)
        self.last_rpm = np.float64(rpm)
        self.last_ang_v_norm = _norm(self.ang_v)    # test instrumentation (conditioning of obs[9:12]), before any reset
        self.rpy = self._euler(self.quat)
        q = self.quat       # distance of Bullet's gimbal-branch test |sarg| >= 0.99999 to its threshold (changes obs[3:6] and the forward vector)
        self.gimbal_margin = np.abs(np.abs(-2.0 * (q[:, 0] * q[:, 2] - q[:, 3] * q[:, 1])) - 0.99999)
        obs = self._obs(self.pos, self.rpy, self.vel, self.ang_v, self.dist)
        # ---- _computeReward (dyn_oracle._reward_waypoint)
        crash_v, final_v, capture_v, cap_orient_w, progress_w, orient_w, smooth = self.rw
        coll = self._collided(self.idx)
        self._m(np.where(coll, np.inf, self.dist - self.threshold))
        captured = ~coll & (self.dist <= self.threshold)
        idx2 = self.idx + captured
        fin = captured & (idx2 == T)
        tgt = self.targets[np.minimum(idx2, T - 1)]
        # orientation_reward
        fwd = np.stack([np.cos(self.rpy[:, 2]) * np.cos(self.rpy[:, 1]), np.sin(self.rpy[:, 2]) * np.cos(self.rpy[:, 1]),
                        np.sin(self.rpy[:, 1])], axis=1)
        d2t = tgt - self.pos
        with np.errstate(all="ignore"):
            d2t = d2t / _norm(d2t)[:, None]
            angle = np.arccos(np.clip(np.sum(fwd * d2t, axis=1), -1.0, 1.0))
            orient = np.where(angle > np.radians(10), -1.0, 0.0)
        # the capture / final rewards stay np.float32 in the reference (np.float32(0.0) + python numbers, PBDroneEnv.py:537-552):
        # (75 - 5) / 25 is the float32 quotient 2.8f
        r_cap = np.float32(capture_v) + ((orient * cap_orient_w).astype(np.float32) if cap_orient_w else np.float32(0.0))
        r_cap = np.float64(np.float32(r_cap) / np.float32(25))
        r = np.exp(-2 * self.dist) * 3
        r = r + np.where(self.just_found, 0.0, (self.prev_dist - self.dist) * progress_w)
        if orient_w:
            r = r + orient * orient_w
        uses_orient = (~coll & ~fin) & ((captured & bool(cap_orient_w)) | (~captured & bool(orient_w)))
        self._m(np.where(uses_orient, angle - np.radians(10), np.inf), "rew_margin")
        if smooth is not None:
            lin = _norm(self.current_vel - self.prev_vel)
            ang = _norm(self.current_ang_v - self.prev_ang_v)
            r = r + np.where(lin > smooth[0], -lin, 0.0) + np.where(ang > smooth[1], -ang, 0.0)
            shaped = ~coll & ~captured
            self._m(np.where(shaped, lin - smooth[0], np.inf), "rew_margin")
            self._m(np.where(shaped, ang - smooth[1], np.inf), "rew_margin")
        reward = np.where(coll, crash_v, np.where(fin, final_v / 25, np.where(captured, r_cap, r / 25)))
        just_found = np.where(coll | fin, self.just_found, captured)
        self.prev_dist = np.where(coll, self.prev_dist, self.dist)
        # ---- _computeTerminated (after the reward, with the possibly advanced index), _computeTruncated, info
        advanced = captured & ~fin
        tube = np.zeros(N, bool)
        if advanced.any() and not self.circle and self.cylinder:
            keep = self.margin.copy()
            tube = self._collided(idx2) & advanced                      # only the segment of the tube depends on the index
            self.margin = np.where(advanced, self.margin, keep)
        terminated = coll | fin | tube
        truncated = self.max_steps <= self.steps
        self.idx = idx2
        self.just_found = just_found
        found = self.idx.copy()
        # ---- _update_state_post_step, skipped when terminated
        nt = ~terminated
        self.steps = self.steps + nt
        self.cur_pos = np.where(nt[:, None], self.pos, self.cur_pos)
        self.prev_vel = np.where(nt[:, None], self.current_vel, self.prev_vel)
        self.prev_ang_v = np.where(nt[:, None], self.current_ang_v, self.prev_ang_v)
        self.current_vel = np.where(nt[:, None], self.vel, self.current_vel)
        self.current_ang_v = np.where(nt[:, None], self.ang_v, self.current_ang_v)
        new_dist = _norm(self.targets[np.minimum(self.idx, T - 1)] - self.cur_pos)
        self.dist = np.where(nt, new_dist, self.dist)
        # ---- wrappers between the env and Monitor (dyn_oracle.OracleWorker.step): observation statistics, reward clip and normalisation
        all_envs = np.ones(N, bool)
        if self.normalize_obs:
            obs = self._normalize_obs(obs, all_envs)
        if self.clip_reward > 0:
            reward = np.clip(reward, -self.clip_reward, self.clip_reward)
        if self.normalize_reward:
            self.returns = self.returns * self.reward_gamma + reward
            self._rms_update(self.returns, self.ret_mean, self.ret_var, self.ret_count, all_envs)
            reward = reward / np.sqrt(self.ret_var + 1e-8)
            self.returns = np.where(terminated | truncated, 0.0, self.returns)
        # ---- Monitor + worker auto-reset
        self.ep_return = self.ep_return + reward
        self.ep_len = self.ep_len + 1
        done = terminated | truncated
        bits = (terminated.astype(np.uint8) | (truncated.astype(np.uint8) << 1))
        terminal_obs = obs.copy()
        ep_r, ep_l = self.ep_return.copy(), self.ep_len.copy()
        if done.any():
            d = done
            # BaseAviary.reset -> _housekeeping; the reset obs is computed BEFORE the distances are reset
            self.pos[d], self.quat[d], self.rpy[d] = self.INIT_XYZ, self.q0, O.bullet_euler_from_quaternion(self.q0)
            self.vel[d], self.ang_v[d], self.rpy_rates[d] = 0.0, 0.0, 0.0
            reset_obs = self._obs(self.pos, self.rpy, self.vel, self.ang_v, self.dist)
            if self.normalize_obs:                                   # NormalizeObservation.reset normalises (and counts) the reset obs too
                reset_obs = self._normalize_obs(reset_obs, d)
            obs = np.where(d[:, None], reset_obs, obs)
            dnew = _norm(self.cur_pos - self.targets[0])               # stale _current_position
            self.dist = np.where(d, dnew, self.dist)
            self.prev_dist = np.where(d, dnew, self.prev_dist)
            self.idx[d], self.steps[d], self.just_found[d] = 0, 0, False
            for a in (self.prev_vel, self.prev_ang_v, self.current_vel, self.current_ang_v):
                a[d] = 0.0
            self.ep_return[d], self.ep_len[d] = 0.0, 0
            self.last_rpm[d] = 0.0
            self.last_clipped[d] = 0.0        # _housekeeping zeroes last_clipped_action (zeros are exact in either dtype)
        return obs, reward, bits, found, terminal_obs, ep_r, ep_l

    def reset_obs(self):
        """VecEnv.reset() right after construction (call once: with normalize_obs it updates the statistics like a reset does)."""
        o = self._obs(self.pos, self.rpy, self.vel, self.ang_v, self.dist)
        return self._normalize_obs(o, np.ones(self.N, bool)) if self.normalize_obs else o

    def state(self):
        """The row-major arrays ``dn_set_state`` takes."""
        return dict(pos=self.pos, quat=self.quat, vel=self.vel, rpy_rates=self.rpy_rates, ang_v=self.ang_v,
                    prev_vel=self.prev_vel, prev_ang_v=self.prev_ang_v, dist=self.dist, prev_dist=self.prev_dist,
                    target_idx=self.idx.astype(np.int32), steps=self.steps.astype(np.int32),
                    just_found=self.just_found.astype(np.uint8), ep_return=self.ep_return, ep_length=self.ep_len.astype(np.int32),
                    **({"last_rpm_sum": np.sum(np.float64(self.last_clipped), axis=1)} if self.drag else {}))
