"""TEST TOOL: ctypes driver for the host build of the device step logic (see emu.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libdnemu.so")

_KSTEPS, _KJF, _KIDX = 0xFFFFF, 1 << 20, 21


def build():
    srcs = [os.path.join(HERE, "emu.cpp"), os.path.join(HERE, "cuda_shim.h")]
    csrc = os.path.join(HERE, "..", "..", "drl-dronenavigation_b200", "csrc")
    srcs += [os.path.join(csrc, f) for f in ("dn_device.cuh", "dn_host.h", "dn_params.h")]
    if os.path.exists(LIB) and all(os.path.getmtime(s) <= os.path.getmtime(LIB) for s in srcs):
        return LIB
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
                    "-include", os.path.join(HERE, "cuda_shim.h"), "-o", LIB, os.path.join(HERE, "emu.cpp")], check=True)
    return LIB


class HostEmuEnv:
    """Same construction arguments as BatchedDroneEnv (subset), numpy in / numpy out."""

    def __init__(self, num_envs, target_points, threshold=0.3, discount=0.999, max_steps=4096,
                 aviary_dim=(-1, -1, 0, 1, 1, 1), initial_xyzs=None, pyb_freq=240, ctrl_freq=240,
                 act_type=0, cylinder=True, circle=False, include_distance=False, normalize_actions=False,
                 physics=0, reward_id=0, normalize_reward=False, clip_reward=0.0, reward_gamma=0.99,
                 random_spawn=False, seed=0, env_id_offset=0, drone_model=0):
        from drl_dronenavigation_b200 import _lib as L
        self.lib = C.CDLL(build())
        self.lib.emu_create.restype = C.c_void_p
        self.lib.emu_create.argtypes = [C.POINTER(L.dn_config)]
        self.lib.emu_action_to_rpm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong]
        self.lib.emu_action_to_rpm.restype = None
        for name, n in (("emu_step", 7), ("emu_get_planes", 2), ("emu_set_planes", 2), ("emu_set_last_rpm_sum", 2), ("emu_destroy", 1),
                        ("emu_get_pid", 2), ("emu_set_pid", 2)):
            getattr(self.lib, name).argtypes = [C.c_void_p] * n
            getattr(self.lib, name).restype = None
        targets = np.ascontiguousarray(np.array(target_points, dtype=np.float64).reshape(-1, 3))
        c = L.dn_config()
        c.abi_version, c.num_envs = L.DN_ABI_VERSION, num_envs
        c.pyb_freq, c.ctrl_freq, c.act_type = pyb_freq, ctrl_freq, act_type
        c.drone_model = drone_model
        c.normalize_actions, c.physics, c.reward_id = int(normalize_actions), physics, reward_id
        c.include_distance, c.cylinder, c.circle, c.max_steps = int(include_distance), int(cylinder), int(circle), max_steps
        c.threshold, c.discount = threshold, discount
        c.spawn_mode, c.seed, c.env_id_offset = ({False: 0, True: 1, "line": 1, "midpoint": 2}[random_spawn]), seed, env_id_offset
        c.normalize_reward, c.clip_reward, c.reward_gamma = int(normalize_reward), float(clip_reward), float(reward_gamma)
        c.aviary_dim = (C.c_double * 6)(*[float(v) for v in aviary_dim])
        c.init_xyz = (C.c_double * 3)(*np.array(initial_xyzs, dtype=np.float64).reshape(-1)[:3])
        c.init_rpy = (C.c_double * 3)(0, 0, 0)
        c.num_targets = targets.shape[0]
        c.targets = targets.ctypes.data_as(C.POINTER(C.c_double))
        self._keep = (targets, c)
        self.h = self.lib.emu_create(C.byref(c))
        assert self.h
        self.num_envs, self.obs_dim = num_envs, 13 if include_distance else 12
        self.uses_drag = bool(physics & 1)
        self._optional = {"pid": act_type >= 3}
        N, D = num_envs, self.obs_dim
        self.obs = np.zeros((N, D), np.float32)
        self.reward = np.zeros(N, np.float32)
        self.done = np.zeros(N, np.uint8)
        self.terminal_obs = np.zeros((N, D), np.float32)
        self.found_targets = np.zeros(N, np.int32)

    def _p(self, a):
        return a.ctypes.data_as(C.c_void_p)

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.float32)
        self.lib.emu_step(self.h, self._p(a), self._p(self.obs), self._p(self.reward), self._p(self.done),
                          self._p(self.terminal_obs), self._p(self.found_targets))
        return self.obs.copy(), self.reward.copy(), self.done.copy(), self.found_targets.copy()

    def action_to_rpm(self, a):
        a = np.ascontiguousarray(a, dtype=np.float32)
        out = np.empty_like(a)
        self.lib.emu_action_to_rpm(self.h, self._p(a), self._p(out), a.size)
        return out

    def _planes(self):
        pl = np.zeros((7, self.num_envs, 4), np.float32)
        self.lib.emu_get_planes(self.h, self._p(pl))
        return pl

    def get_state(self):
        pl = self._planes()
        bits = pl[4, :, 3].view(np.uint32)
        return dict(pos=pl[0, :, :3].copy(), dist=pl[0, :, 3].copy(), quat=pl[1].copy(), vel=pl[2, :, :3].copy(),
                    prev_dist=pl[2, :, 3].copy(), rpy_rates=pl[3, :, :3].copy(), ep_return=pl[3, :, 3].copy(),
                    ang_v=pl[4, :, :3].copy(), steps=(bits & _KSTEPS).astype(np.int32),
                    just_found=((bits & _KJF) != 0).astype(np.uint8), target_idx=(bits >> _KIDX).astype(np.int32),
                    prev_vel=pl[5, :, :3].copy(), ep_length=pl[5, :, 3].view(np.int32).copy(),
                    prev_ang_v=pl[6, :, :3].copy(), episode_count=pl[6, :, 3].view(np.int32).copy(), pid=self.get_pid())

    def get_pid(self):
        out = np.zeros((self.num_envs, 9), np.float32)
        self.lib.emu_get_pid(self.h, self._p(out))
        return out

    def set_state(self, st):
        pl = self._planes()
        for k, (p, sl) in dict(pos=(0, slice(0, 3)), quat=(1, slice(0, 4)), vel=(2, slice(0, 3)), rpy_rates=(3, slice(0, 3)),
                               ang_v=(4, slice(0, 3)), prev_vel=(5, slice(0, 3)), prev_ang_v=(6, slice(0, 3))).items():
            if k in st:
                pl[p, :, sl] = st[k]
        for k, p in dict(dist=0, prev_dist=2, ep_return=3).items():
            if k in st:
                pl[p, :, 3] = st[k]
        cur = self.get_state()
        g = lambda k: np.asarray(st.get(k, cur[k])).astype(np.uint32)
        pl[4, :, 3] = ((g("target_idx") << _KIDX) | (g("just_found") * _KJF) | (g("steps") & _KSTEPS)).astype(np.uint32).view(np.float32)
        if "ep_length" in st:
            pl[5, :, 3] = np.asarray(st["ep_length"], np.int32).view(np.float32)
        self.lib.emu_set_planes(self.h, self._p(np.ascontiguousarray(pl)))
        if "pid" in st:
            self.lib.emu_set_pid(self.h, self._p(np.ascontiguousarray(st["pid"], dtype=np.float32)))
        if self.uses_drag and "last_rpm_sum" in st:
            self.lib.emu_set_last_rpm_sum(self.h, self._p(np.ascontiguousarray(st["last_rpm_sum"], dtype=np.float32)))

    def reset(self):
        """explicit reset of fresh envs only (the emulator has no reset kernel): constructor state"""
        st = self.get_state()
        assert (st["steps"] == 0).all()
        o = np.zeros((self.num_envs, self.obs_dim), np.float32)
        return o

    def close(self):
        self.lib.emu_destroy(self.h)
