"""gymnasium-style single-environment facade: ``PBDroneEnv``.

Same constructor keywords and ``reset`` / ``step`` return conventions as
Sol/Model/Environments/PBDroneEnv.py:41-65,171-199,609-665, backed by a one-env
``BatchedDroneEnv`` (so a single ``step`` is still a CUDA kernel launch; this class is
for the reference's ``run_test`` / ``test_saved`` style loops, not for throughput).
Unlike the vectorised env it does NOT auto-reset: after a terminal step the caller calls
``reset()``, exactly like the reference env.
"""
from __future__ import annotations

import numpy as np
import torch

from .batched_env import BatchedDroneEnv
from .constants import CF2X
from .enums import ActionType, DroneModel, ObservationType, Physics
from .vec_env import action_space, observation_space


class PBDroneEnv:
    metadata = {"render_modes": []}

    def __init__(self, target_points, threshold, discount, max_steps, aviary_dim,
                 save_folder=None, drone_model: DroneModel = DroneModel.CF2X, initial_xyzs=None,
                 initial_rpys=None, physics: Physics = Physics.PYB, pyb_freq: int = 240, ctrl_freq: int = 240,
                 gui=False, record=False, obs: ObservationType = ObservationType.KIN,
                 act: ActionType = ActionType.THRUST, vision_attributes=False, user_debug_gui=False,
                 obstacles=False, random_spawn=False, cylinder=True, circle=False, include_target=False,
                 include_distance=False, normalize_actions=False, collect_rollouts=False, device=None, seed: int = 0):
        if gui or record or vision_attributes or obstacles:
            raise NotImplementedError("rendering / vision / obstacles are outside the CUDA hot path")
        self._core = BatchedDroneEnv(1, target_points, threshold=threshold, discount=discount, max_steps=max_steps,
                                     aviary_dim=aviary_dim, initial_xyzs=initial_xyzs, initial_rpys=initial_rpys,
                                     drone_model=drone_model, physics=physics, pyb_freq=pyb_freq, ctrl_freq=ctrl_freq,
                                     obs=obs, act=act, cylinder=cylinder, circle=circle,
                                     include_distance=include_distance, normalize_actions=normalize_actions,
                                     normalize_obs=False, device=device,
                                     # the reference's make_env passes False (PBDroneSimulator.py:166); True selects the
                                     # Philox version of its commented-out spawn around a target-pair line (PBDroneEnv.py:622-629)
                                     random_spawn=random_spawn, seed=seed)
        self.ACT_TYPE, self.OBS_TYPE, self.PHYSICS = act, obs, physics
        self.normalize_actions, self.include_distance = normalize_actions, include_distance
        self.action_space = action_space(normalize_actions)
        self.observation_space = observation_space(include_distance)
        self.physical_action_bounds = CF2X.physical_action_bounds()
        self.G, self.CTRL_FREQ, self.PYB_FREQ = CF2X.G, ctrl_freq, pyb_freq
        self.INIT_XYZS, self.INIT_RPYS = self._core.INIT_XYZS, self._core.INIT_RPYS
        self._actions = torch.zeros(1, 4, dtype=torch.float32, device=self._core.device)
        self._auto_reset_done = False  # the kernel already reset the env on the last (terminal) step
        self.last_clipped_action = np.zeros((1, 4))
        self._refresh()

    # kinematics the reference's manager pokes (PBDroneSimulator.py:406-407,532-533,775)
    def _refresh(self):
        st = {k: v.cpu().numpy() for k, v in self._core.get_state().items()}
        self.pos, self.quat, self.vel, self.ang_v = st["pos"], st["quat"], st["vel"], st["ang_v"]
        x, y, z, w = [float(v) for v in self.quat[0]]
        sarg = -2.0 * (x * z - w * y)
        if sarg <= -0.99999:
            rpy = (0.0, -0.5 * np.pi, 2 * np.arctan2(x, -y))
        elif sarg >= 0.99999:
            rpy = (0.0, 0.5 * np.pi, 2 * np.arctan2(-x, y))
        else:
            rpy = (np.arctan2(2 * (y * z + w * x), w * w - x * x - y * y + z * z), np.arcsin(sarg),
                   np.arctan2(2 * (x * y + w * z), w * w + x * x - y * y - z * z))
        self.rpy = np.array([rpy])
        self._current_target_index = int(st["target_idx"][0])
        self._steps = int(st["steps"][0])
        self._distance_to_target = float(st["dist"][0])

    def reset(self, seed: int = None, options: dict = None):
        if self._auto_reset_done:
            # the fused step already performed this reset (same state, same observation)
            obs = self._core.obs.cpu().numpy()[0].copy()
        else:
            obs = self._core.reset().cpu().numpy()[0].copy()
        self._auto_reset_done = False
        self.last_clipped_action = np.zeros((1, 4))
        info = {"found_targets": 0}
        self._refresh()
        return obs, info

    def step(self, action):
        a = np.asarray(action, dtype=np.float32).reshape(1, 4)
        self._actions.copy_(torch.from_numpy(a))
        c = self._core
        c.step(self._actions)
        if self.ACT_TYPE in (ActionType.PID, ActionType.VEL, ActionType.ONE_D_PID):
            # the RPMs of the PID family depend on the controller state inside the kernel; not exported per step
            self.last_clipped_action = np.full((1, 4), np.nan)
        else:
            self.last_clipped_action = c.action_to_rpm(self._actions).cpu().numpy().astype(np.float64)   # BaseAviary.py:442
        bits = int(c.done.cpu()[0])
        terminated, truncated = bool(bits & 1), bool(bits & 2)
        reward = float(c.reward.cpu()[0])
        info = {"found_targets": int(c.found_targets.cpu()[0])}
        # on done the kernel has already auto-reset (VecEnv semantics): hand back the terminal obs
        obs = (c.terminal_obs if bits else c.obs).cpu().numpy()[0].copy()
        self._auto_reset_done = bool(bits)
        self._refresh()
        return obs, reward, terminated, truncated, info

    def _getDroneStateVector(self, nth_drone: int = 0):
        """BaseAviary._getDroneStateVector (BaseAviary.py:623-643): the 20-vector
        [pos3, quat4, rpy3, vel3, ang_v3, last_clipped_action4] of the (single) drone; the last four entries are the
        motor RPMs of the most recent step (zeros after a reset, BaseAviary.py:545)."""
        if nth_drone != 0:
            raise IndexError("single-drone environment")
        return np.hstack([self.pos[0], self.quat[0], self.rpy[0], self.vel[0], self.ang_v[0], self.last_clipped_action[0]]).reshape(20,)

    def close(self):
        self._core.close()

    def getPyBulletClient(self):
        return -1
