"""TEST TOOL: the host build of the device step logic (tests/host_emu) behind BatchedDroneEnv's tensor interface, on the
CPU -- lets the CPU suite drive the manager (PBDroneSimulator) and the learners end to end through the real step logic.
Never imported by the package."""
import numpy as np
import torch

from tests.host_emu import HostEmuEnv


class EmuTorchEnv:
    def __init__(self, num_envs, target_points, threshold=0.3, discount=0.999, max_steps=4096, aviary_dim=(-1, -1, 0, 1, 1, 1),
                 initial_xyzs=None, pyb_freq=240, ctrl_freq=240, cylinder=True, circle=False, include_distance=True,
                 normalize_actions=True, reward_id=0, normalize_reward=False, clip_reward=0.0, **ignored):
        self.emu = HostEmuEnv(num_envs, target_points, threshold=threshold, discount=discount, max_steps=max_steps, aviary_dim=aviary_dim,
                              initial_xyzs=initial_xyzs, pyb_freq=pyb_freq, ctrl_freq=ctrl_freq, cylinder=cylinder, circle=circle,
                              include_distance=include_distance, normalize_actions=normalize_actions, reward_id=reward_id,
                              normalize_reward=normalize_reward, clip_reward=clip_reward)
        self.device, self.num_envs, self.obs_dim = torch.device("cpu"), num_envs, self.emu.obs_dim
        self.num_targets = len(np.asarray(target_points).reshape(-1, 3))
        init = np.asarray(initial_xyzs, np.float64).reshape(-1)[:3]
        dim = [float(v) for v in aviary_dim]
        d0 = np.linalg.norm(init - np.asarray(target_points, np.float64).reshape(-1, 3)[0]) / max(abs(dim[0]) + dim[3], abs(dim[1]) + dim[4], dim[5])
        row = [init[0] / dim[3], init[1] / dim[4], init[2] / dim[5]] + [0.0] * 9 + ([d0] if include_distance else [])
        self._obs0 = torch.tensor(np.tile(np.array(row, np.float32), (num_envs, 1)))     # the constructor-state observation
        self.terminal_obs = torch.zeros(num_envs, self.obs_dim)
        self._ret, self._len = np.zeros(num_envs), np.zeros(num_envs, np.int64)
        self._stats = dict(return_sum=0.0, length_sum=0, episodes=0, successes=0, found_targets=0, crashes=0, truncations=0)

    def reset(self, mask=None):
        return self._obs0.clone()

    def step(self, actions):
        o, r, d, f = self.emu.step(actions.detach().cpu().numpy())
        self.terminal_obs = torch.from_numpy(self.emu.terminal_obs.copy())
        self._ret += r
        self._len += 1
        done = d != 0
        if done.any():        # the Monitor statistics dn_episode_stats keeps on the device
            s = self._stats
            succ = done & (f == self.num_targets)
            s["return_sum"] += float(self._ret[done].sum()); s["length_sum"] += int(self._len[done].sum()); s["episodes"] += int(done.sum())
            s["successes"] += int(succ.sum()); s["found_targets"] += int(f[done].sum())
            s["crashes"] += int((done & ((d & 1) != 0) & ~succ).sum()); s["truncations"] += int((d == 2).sum())
            self._ret[done], self._len[done] = 0.0, 0
        return torch.from_numpy(o), torch.from_numpy(r), torch.from_numpy(d), torch.from_numpy(f)

    def episode_stats(self, clear=False):
        out = dict(self._stats)
        if clear:
            for k in self._stats:
                self._stats[k] = 0 if k != "return_sum" else 0.0
        return out

    def close(self):
        self.emu.close()
