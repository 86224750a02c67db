"""Torch-native SAC on the device-resident environment (BASELINE config 4).

Hyper-parameters are the SAC branch of ``PBDroneSimulator.setup_agent`` (Sol/Model/PBDroneSimulator.py:290-331:
qf [256, 256, 128], pi [256, 256], ReLU, batch 1024, buffer 1 048 576, learning_starts 8192, train_freq 3,
gradient_steps 5, tau 0.005, lr 2.5e-4, gamma 0.99, ent_coef "auto", target_entropy "auto"); the update is the
maths of SB3's ``SAC.train`` that the reference relies on (squashed diagonal Gaussian actor with log-std clamped to
[-20, 2], twin critics, entropy-coefficient auto-tuning towards -|A|, Polyak targets every gradient step, time-limit
truncations not treated as terminal, next observation of a finished episode = its terminal observation).
What changes is where the data lives: the replay buffer is a set of device tensors filled straight from the fused
step kernel's outputs (no numpy hop, no per-env Python), and under ``torchrun`` every gradient step all-reduces two
flat buckets (critics + log-alpha, actor) over NCCL -- the environment shards need no collective.
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.distributed as dist
from torch import nn

LOG_STD_MIN, LOG_STD_MAX = -20.0, 2.0      # SB3 sac/policies.py


@dataclass
class SACConfig:
    """Defaults = PBDroneSimulator.setup_agent's SAC branch (PBDroneSimulator.py:304-327)."""
    learning_starts: int = 8192
    train_freq: int = 3            # env.step calls (x num_envs transitions) between update phases
    gradient_steps: int = 5
    batch_size: int = 1024
    tau: float = 0.005
    target_update_interval: int = 1
    buffer_size: int = 1_048_576   # transitions in total (SB3 divides by n_envs)
    learning_rate: float = 2.5e-4
    gamma: float = 0.99
    pi_arch: tuple = (256, 256)
    qf_arch: tuple = (256, 256, 128)
    n_critics: int = 2
    ent_coef_init: float = 1.0     # ent_coef="auto" -> log_ent_coef = log(1.0)
    target_entropy: Optional[float] = None   # "auto" -> -|A|
    seed: int = 42
    matmul_precision: str = "tf32"
    cuda_graph: bool = True        # replay each gradient step from three captured CUDA graphs (CUDA devices only)


def _relu_mlp(sizes, out_dim=None):
    layers = []
    for a, b in zip(sizes[:-1], sizes[1:]):
        layers += [nn.Linear(a, b), nn.ReLU()]
    if out_dim is not None:
        layers.append(nn.Linear(sizes[-1], out_dim))
    return nn.Sequential(*layers)


class Actor(nn.Module):
    """SB3 sac.policies.Actor: latent MLP -> (mu, log_std) heads, tanh-squashed Gaussian."""

    def __init__(self, obs_dim, act_dim, arch):
        super().__init__()
        self.latent = _relu_mlp((obs_dim,) + tuple(arch))
        self.mu = nn.Linear(arch[-1], act_dim)
        self.log_std = nn.Linear(arch[-1], act_dim)

    def forward(self, obs, generator=None, deterministic=False, eps=None):
        h = self.latent(obs)
        mu, log_std = self.mu(h), self.log_std(h).clamp(LOG_STD_MIN, LOG_STD_MAX)
        if deterministic:
            return torch.tanh(mu), None
        std = log_std.exp()
        if eps is None:            # (the graph-replayed update passes pre-drawn noise: no RNG inside a captured graph)
            eps = torch.randn(mu.shape, device=mu.device, dtype=mu.dtype, generator=generator)
        g = mu + std * eps
        a = torch.tanh(g)
        # Normal(mu, std).log_prob(g).sum - sum log(1 - tanh(g)^2 + eps)   (SquashedDiagGaussianDistribution, epsilon 1e-6)
        logp = (-0.5 * eps.pow(2) - log_std - 0.5 * math.log(2 * math.pi)).sum(-1) - torch.log(1 - a.pow(2) + 1e-6).sum(-1)
        return a, logp


class Critics(nn.Module):
    """n_critics independent Q(s, a) MLPs (SB3 ContinuousCritic, share_features_extractor=False)."""

    def __init__(self, obs_dim, act_dim, arch, n):
        super().__init__()
        self.qs = nn.ModuleList([_relu_mlp((obs_dim + act_dim,) + tuple(arch), 1) for _ in range(n)])

    def forward(self, obs, act):
        x = torch.cat([obs, act], dim=-1)
        return [q(x).squeeze(-1) for q in self.qs]


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class _FlatGrads:
    """All gradients of a parameter list as views into ONE tensor (one all-reduce per optimiser step)."""

    def __init__(self, params, device):
        self.params = list(params)
        self.flat = torch.zeros(sum(p.numel() for p in self.params), device=device)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.calls = 0

    def zero(self):
        self.flat.zero_()

    def allreduce(self):
        w = _world()
        if w > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(w)
            self.calls += 1


class ReplayBuffer:
    """Device-resident circular buffer, [capacity_steps, N, ...] (SB3 ReplayBuffer with n_envs = N,
    handle_timeout_termination=True, optimize_memory_usage=False)."""

    def __init__(self, capacity_transitions: int, num_envs: int, obs_dim: int, act_dim: int, device):
        self.cap = max(capacity_transitions // num_envs, 1)
        self.N, self.pos, self.full = num_envs, 0, False
        f = dict(dtype=torch.float32, device=device)
        self.obs = torch.empty(self.cap, num_envs, obs_dim, **f)
        self.next_obs = torch.empty(self.cap, num_envs, obs_dim, **f)
        self.act = torch.empty(self.cap, num_envs, act_dim, **f)
        self.rew = torch.empty(self.cap, num_envs, **f)
        self.done = torch.empty(self.cap, num_envs, **f)        # 1 only for true terminations (timeouts excluded)

    def add(self, obs, next_obs, act, rew, done_bits, terminal_obs):
        p = self.pos
        finished = done_bits != 0
        self.obs[p].copy_(obs)
        # the VecEnv returns the RESET observation on done; the transition's successor is the terminal observation
        self.next_obs[p].copy_(torch.where(finished.unsqueeze(-1), terminal_obs, next_obs))
        self.act[p].copy_(act)
        self.rew[p].copy_(rew)
        self.done[p].copy_(((done_bits & 1) != 0).to(torch.float32))      # DN_DONE_TERMINATED; truncated-only -> bootstrap
        self.pos = (p + 1) % self.cap
        self.full = self.full or self.pos == 0

    def __len__(self):
        return (self.cap if self.full else self.pos) * self.N

    # ---- SaveReplayBufferCallback / load_replay_buffer (Sol/Utilities/Callbacks.py:13-39; PBDroneSimulator.py:357,998-1017) ----
    def save(self, path: str) -> str:
        """``model.save_replay_buffer(path)``: SB3 pickles its ReplayBuffer object; here a dict with that object's attribute
        names and array layouts ([buffer_size, n_envs, ...]) is pickled, which :meth:`load` -- and anything that reads the
        attributes of an SB3 buffer -- understands."""
        import pickle
        n = self.cap if self.full else self.pos
        arr = lambda t: t[:n].detach().cpu().numpy() if not self.full else t.detach().cpu().numpy()
        with open(path, "wb") as f:
            pickle.dump({"observations": arr(self.obs), "next_observations": arr(self.next_obs), "actions": arr(self.act),
                         "rewards": arr(self.rew), "dones": arr(self.done), "pos": self.pos, "full": self.full,
                         "buffer_size": self.cap, "n_envs": self.N, "written_by": "drl_dronenavigation_b200"}, f, protocol=4)
        return path

    def load(self, path: str) -> int:
        """``model.load_replay_buffer(path)``: accepts the dict written by :meth:`save` or a pickled SB3 ``ReplayBuffer`` (when
        SB3 is importable); returns the number of transitions restored.  n_envs must match; a larger saved buffer is truncated
        to the most recent ``capacity`` steps."""
        import pickle
        with open(path, "rb") as f:
            obj = pickle.load(f)
        get = (lambda k: obj[k]) if isinstance(obj, dict) else (lambda k: getattr(obj, k))
        n_envs, pos, full = int(get("n_envs")), int(get("pos")), bool(get("full"))
        if n_envs != self.N:
            raise ValueError(f"replay buffer was saved with n_envs={n_envs}, this trainer has {self.N}")
        def ordered(a):                      # oldest -> newest
            a = torch.as_tensor(a, dtype=torch.float32)
            if full and a.shape[0] > pos:
                a = torch.cat([a[pos:], a[:pos]])
            elif not full:
                a = a[:pos]
            return a[-self.cap:]
        fields = {"obs": "observations", "next_obs": "next_observations", "act": "actions", "rew": "rewards", "done": "dones"}
        # SB3 samples dones * (1 - timeouts) (ReplayBuffer._get_samples with handle_timeout_termination): a time-limit truncation is
        # not a terminal state and must keep its bootstrap.  This buffer stores terminated-only flags in `done`, so an SB3 object
        # that carries `timeouts` is converted on the way in.
        timeouts = None
        try:
            timeouts = get("timeouts")
        except (KeyError, AttributeError):
            pass
        n = 0
        for ours, theirs in fields.items():
            a = ordered(get(theirs))
            if ours == "done" and timeouts is not None:
                a = a * (1.0 - ordered(timeouts).reshape(a.shape))
            dst = getattr(self, ours)
            a = a.reshape((a.shape[0],) + tuple(dst.shape[1:]))
            n = a.shape[0]
            dst[:n].copy_(a.to(dst.device))
        self.full = n == self.cap
        self.pos = 0 if self.full else n
        return n * self.N

    def sample(self, batch_size: int, generator=None):
        n = len(self)
        idx = torch.randint(0, n, (batch_size,), device=self.obs.device, generator=generator)
        flat = lambda t: t.reshape(self.cap * self.N, *t.shape[2:])
        return flat(self.obs)[idx], flat(self.act)[idx], flat(self.rew)[idx], flat(self.next_obs)[idx], flat(self.done)[idx]


class SACLearner:
    """Actor, twin critics + targets, log-alpha and the SAC update; independent of the environment."""

    def __init__(self, obs_dim: int, act_dim: int, cfg: SACConfig = SACConfig(), device="cpu"):
        self.cfg, self.device = cfg, torch.device(device)
        torch.manual_seed(cfg.seed)                         # same seed on every rank -> identical initial parameters
        self.actor = Actor(obs_dim, act_dim, cfg.pi_arch).to(self.device)
        self.critic = Critics(obs_dim, act_dim, cfg.qf_arch, cfg.n_critics).to(self.device)
        self.critic_target = Critics(obs_dim, act_dim, cfg.qf_arch, cfg.n_critics).to(self.device)
        self.critic_target.load_state_dict(self.critic.state_dict())
        for p in self.critic_target.parameters():
            p.requires_grad_(False)
        self.log_ent_coef = torch.log(torch.ones(1, device=self.device) * cfg.ent_coef_init).requires_grad_(True)
        self.target_entropy = float(-act_dim) if cfg.target_entropy is None else float(cfg.target_entropy)
        lr = cfg.learning_rate
        # (the Polyak decision is baked into the captured graph, so replay is only used with target_update_interval == 1)
        self.use_graph = bool(cfg.cuda_graph) and self.device.type == "cuda" and cfg.target_update_interval == 1
        self._graphs = None
        self.actor_opt = torch.optim.Adam(self.actor.parameters(), lr=lr, capturable=self.use_graph)
        self.critic_opt = torch.optim.Adam(self.critic.parameters(), lr=lr, capturable=self.use_graph)
        self.ent_opt = torch.optim.Adam([self.log_ent_coef], lr=lr, capturable=self.use_graph)
        # bucket 1: critics + log-alpha (their losses do not depend on each other's step); bucket 2: actor
        self.g_critic = _FlatGrads(list(self.critic.parameters()) + [self.log_ent_coef], self.device)
        self.g_actor = _FlatGrads(self.actor.parameters(), self.device)
        self.n_updates = 0
        if self.device.type == "cuda":
            torch.backends.cuda.matmul.allow_tf32 = (cfg.matmul_precision == "tf32")

    @torch.no_grad()
    def act(self, obs, generator=None, deterministic=False):
        return self.actor(obs, generator=generator, deterministic=deterministic)[0]

    # ---- one gradient step of SB3's SAC.train, in three stages around the two gradient all-reduces -----------------
    def _stage_critic(self, obs, act, rew, next_obs, done, eps_pi, eps_next, out):
        """actor forward, entropy-coefficient loss, TD target, critic loss; backward into the critics + log-alpha bucket."""
        cfg = self.cfg
        actions_pi, log_prob = self.actor(obs, eps=eps_pi)
        ent_coef = self.log_ent_coef.detach().exp()
        ent_loss = -(self.log_ent_coef * (log_prob + self.target_entropy).detach()).mean()
        with torch.no_grad():
            next_a, next_logp = self.actor(next_obs, eps=eps_next)
            next_q = torch.stack(self.critic_target(next_obs, next_a), 0).min(0).values - ent_coef * next_logp
            target_q = rew + (1.0 - done) * cfg.gamma * next_q
        cur_q = self.critic(obs, act)
        critic_loss = 0.5 * sum(torch.nn.functional.mse_loss(q, target_q) for q in cur_q)
        self.g_critic.zero()
        (critic_loss + ent_loss).backward(inputs=self.g_critic.params)
        out[0].copy_(critic_loss.detach()); out[2].copy_(ent_coef.squeeze(0)); out[3].copy_(ent_loss.detach())
        return actions_pi, log_prob, ent_coef

    def _stage_actor(self, obs, actions_pi, log_prob, ent_coef, out):
        """optimiser steps of log-alpha and the critics, then the actor loss against the UPDATED critics (as in SB3)."""
        self.ent_opt.step()
        self.critic_opt.step()
        q_pi = torch.stack(self.critic(obs, actions_pi), 0).min(0).values
        actor_loss = (ent_coef * log_prob - q_pi).mean()
        self.g_actor.zero()
        actor_loss.backward(inputs=self.g_actor.params)
        out[1].copy_(actor_loss.detach())

    def _stage_finish(self):
        self.actor_opt.step()
        if (self.n_updates + 1) % self.cfg.target_update_interval == 0:
            with torch.no_grad():                            # polyak_update(critic, critic_target, tau)
                tp, sp = list(self.critic_target.parameters()), list(self.critic.parameters())
                torch._foreach_mul_(tp, 1.0 - self.cfg.tau)
                torch._foreach_add_(tp, sp, alpha=self.cfg.tau)

    def _capture(self, B, D, A):
        """Three CUDA graphs per gradient step ([actor fwd + critic losses + backward] | all-reduce | [alpha / critic
        steps + actor loss + backward] | all-reduce | [actor step + Polyak]): a 1024-sample SAC step is ~300 small
        kernels, i.e. launch-bound when issued eagerly."""
        dev = self.device
        f = dict(dtype=torch.float32, device=dev)
        st = dict(obs=torch.zeros(B, D, **f), act=torch.zeros(B, A, **f), rew=torch.zeros(B, **f), next_obs=torch.zeros(B, D, **f),
                  done=torch.zeros(B, **f), eps_pi=torch.zeros(B, A, **f), eps_next=torch.zeros(B, A, **f), out=torch.zeros(4, **f))
        nets = list(self.actor.parameters()) + list(self.critic.parameters()) + list(self.critic_target.parameters()) + [self.log_ent_coef]
        saved = [p.detach().clone() for p in nets]
        snap = [[{k: v.clone() for k, v in o.state[p].items() if torch.is_tensor(v)} if p in o.state else None
                 for g in o.param_groups for p in g["params"]] for o in (self.actor_opt, self.critic_opt, self.ent_opt)]
        keep = {}

        def run_all():
            keep["a"], keep["l"], keep["e"] = self._stage_critic(st["obs"], st["act"], st["rew"], st["next_obs"], st["done"],
                                                                 st["eps_pi"], st["eps_next"], st["out"])
            self._stage_actor(st["obs"], keep["a"], keep["l"], keep["e"], st["out"])
            self._stage_finish()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        if hasattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
            # the gradient buckets were created on the current stream, the warm-up runs on a side stream: intended
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        with torch.cuda.stream(side):
            st["obs"].normal_(); st["next_obs"].normal_(); st["act"].uniform_(-1, 1); st["eps_pi"].normal_(); st["eps_next"].normal_()
            for _ in range(3):
                run_all()
        torch.cuda.current_stream(dev).wait_stream(side)
        g1, g2, g3 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        pool = torch.cuda.graph_pool_handle()                # stage 2 consumes tensors stage 1 produced: one shared pool
        with torch.cuda.graph(g1, pool=pool):
            keep["a"], keep["l"], keep["e"] = self._stage_critic(st["obs"], st["act"], st["rew"], st["next_obs"], st["done"],
                                                                 st["eps_pi"], st["eps_next"], st["out"])
        with torch.cuda.graph(g2, pool=pool):
            self._stage_actor(st["obs"], keep["a"], keep["l"], keep["e"], st["out"])
        with torch.cuda.graph(g3, pool=pool):
            self._stage_finish()
        with torch.no_grad():                                # undo the warm-up: parameters, targets, Adam state
            for p, q in zip(nets, saved):
                p.copy_(q)
            for o, sn in zip((self.actor_opt, self.critic_opt, self.ent_opt), snap):
                ps = [p for g in o.param_groups for p in g["params"]]
                for p, old in zip(ps, sn):
                    for k, v in o.state[p].items():
                        if torch.is_tensor(v):
                            v.copy_(old[k]) if old is not None else v.zero_()
        self._graphs = ((B, D, A), g1, g2, g3, st, keep)

    def update(self, batch, generator=None) -> Dict[str, torch.Tensor]:
        """One gradient step of SB3's SAC.train on a sampled batch (obs, act, rew, next_obs, done)."""
        obs, act, rew, next_obs, done = batch
        B, D, A = obs.shape[0], obs.shape[1], act.shape[1]
        if self.use_graph and obs.is_cuda:
            if self._graphs is None or self._graphs[0] != (B, D, A):
                self._capture(B, D, A)
            _, g1, g2, g3, st, _ = self._graphs
            st["obs"].copy_(obs); st["act"].copy_(act); st["rew"].copy_(rew); st["next_obs"].copy_(next_obs); st["done"].copy_(done)
            st["eps_pi"].normal_(generator=generator); st["eps_next"].normal_(generator=generator)
            g1.replay()
            self.g_critic.allreduce()
            g2.replay()
            self.g_actor.allreduce()
            g3.replay()
            out = st["out"].clone()
        else:
            out = torch.zeros(4, device=self.device)
            eps_pi = torch.randn(B, A, device=self.device, generator=generator)
            eps_next = torch.randn(B, A, device=self.device, generator=generator)
            actions_pi, log_prob, ent_coef = self._stage_critic(obs, act, rew, next_obs, done, eps_pi, eps_next, out)
            self.g_critic.allreduce()
            self._stage_actor(obs, actions_pi, log_prob, ent_coef, out)
            self.g_actor.allreduce()
            self._stage_finish()
        self.n_updates += 1
        return {"critic_loss": out[0], "actor_loss": out[1], "ent_coef": out[2], "ent_coef_loss": out[3]}

    def flat_parameters(self) -> torch.Tensor:
        ps = list(self.actor.parameters()) + list(self.critic.parameters()) + [self.log_ent_coef]
        return torch.cat([p.detach().reshape(-1) for p in ps])


class SACTrainer:
    """Off-policy loop on a BatchedDroneEnv shard: train_freq env steps, then gradient_steps updates."""

    def __init__(self, env, cfg: SACConfig = SACConfig()):
        self.env, self.cfg, self.dev = env, cfg, env.device
        self.learner = SACLearner(env.obs_dim, 4, cfg, device=self.dev)
        rank = dist.get_rank() if _world() > 1 else 0
        self.gen = torch.Generator(device=self.dev).manual_seed(cfg.seed + 1000 * (rank + 1))
        self.buffer = ReplayBuffer(cfg.buffer_size, env.num_envs, env.obs_dim, 4, self.dev)
        self.obs = env.reset().clone()
        self.total_steps = 0

    @torch.no_grad()
    def collect(self, n_steps: int):
        env, N = self.env, self.env.num_envs
        for _ in range(n_steps):
            if self.total_steps < self.cfg.learning_starts:      # SB3: uniform random actions before learning_starts
                a = torch.rand(N, 4, device=self.dev, generator=self.gen) * 2 - 1
            else:
                a = self.learner.act(self.obs, generator=self.gen)
            a = a.contiguous()
            obs, rew, done, _ = env.step(a)
            self.buffer.add(self.obs, obs, a, rew, done, env.terminal_obs)
            self.obs = obs.clone()
            self.total_steps += N * _world()

    def train_iteration(self) -> Dict[str, float]:
        cfg = self.cfg
        t0 = time.perf_counter()
        self.collect(cfg.train_freq)
        out = {}
        grad_steps = 0
        if self.total_steps >= cfg.learning_starts and len(self.buffer) >= cfg.batch_size:
            for _ in range(cfg.gradient_steps):
                out = self.learner.update(self.buffer.sample(cfg.batch_size, generator=self.gen), generator=self.gen)
                grad_steps += 1
        if self.dev.type == "cuda":
            torch.cuda.synchronize(self.dev)
        dt = time.perf_counter() - t0
        res = {k: float(v) for k, v in out.items()}
        res.update(samples=cfg.train_freq * self.env.num_envs * _world(), gradient_steps=grad_steps, seconds=dt,
                   sps=cfg.train_freq * self.env.num_envs * _world() / dt)
        return res
