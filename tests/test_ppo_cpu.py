"""Host-side logic of the PPO learner (no GPU): GAE recursion, the update step, and the N>1 path
(one flat-gradient all-reduce per optimiser step + synchronised early stop) on the gloo backend."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from drl_dronenavigation_b200.ppo import ActorCritic, PPOConfig, PPOLearner, compute_gae_torch


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_gae_matches_sb3_recursion():
    rng = np.random.default_rng(0)
    T, N, gamma, lam = 37, 5, 0.99, 0.95
    rew, val = rng.normal(size=(T, N)).astype(np.float32), rng.normal(size=(T, N)).astype(np.float32)
    done = (rng.random((T, N)) < 0.1).astype(np.uint8)
    last = rng.normal(size=N).astype(np.float32)
    # SB3 RolloutBuffer.compute_returns_and_advantage written with episode_starts
    episode_starts = np.zeros((T + 1, N), np.float32)
    episode_starts[1:] = done
    adv = np.zeros((T, N), np.float32)
    last_gae = 0
    for step in reversed(range(T)):
        next_non_terminal = 1.0 - episode_starts[step + 1]
        next_values = last if step == T - 1 else val[step + 1]
        delta = rew[step] + gamma * next_values * next_non_terminal - val[step]
        last_gae = delta + gamma * lam * next_non_terminal * last_gae
        adv[step] = last_gae
    a, r = compute_gae_torch(torch.from_numpy(rew), torch.from_numpy(val), torch.from_numpy(done), torch.from_numpy(last), gamma, lam)
    np.testing.assert_allclose(a.numpy(), adv, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(r.numpy(), adv + val, rtol=1e-5, atol=1e-5)


def test_policy_matches_sb3_defaults():
    cfg = PPOConfig()
    pol = ActorCritic(13, 4, cfg)
    n = sum(p.numel() for p in pol.parameters())
    # 2 x [13->512->512->256] + heads (4, 1) + log_std(4): ~0.80 M parameters = 3.2 MB FP32 bucket
    assert n == 2 * (13 * 512 + 512 + 512 * 512 + 512 + 512 * 256 + 256) + (256 * 4 + 4) + (256 + 1) + 4
    assert torch.allclose(pol.log_std, torch.zeros(4))
    obs = torch.randn(7, 13)
    a, logp, v = pol.act(obs, deterministic=True)
    v2, logp2, ent = pol.evaluate(obs, a)
    assert torch.allclose(logp, logp2) and torch.allclose(v, v2)
    assert torch.allclose(ent, torch.full((7,), 4 * (0.5 + 0.5 * np.log(2 * np.pi))), atol=1e-6)


def _synthetic(rank, B=256, seed=0):
    g = torch.Generator().manual_seed(seed + rank)
    obs = torch.randn(B, 13, generator=g)
    act = torch.randn(B, 4, generator=g).clamp(-1, 1)
    adv = torch.randn(B, generator=g)
    ret = torch.randn(B, generator=g)
    return obs, act, adv, ret


def test_update_improves_surrogate_and_is_deterministic():
    cfg = PPOConfig(batch_size=64, n_epochs=3, target_kl=None)
    outs = []
    for _ in range(2):
        L = PPOLearner(13, 4, cfg)
        obs, act, adv, ret = _synthetic(0)
        with torch.no_grad():
            v, logp, _ = L.policy.evaluate(obs, act)
        L.update(obs, act, logp, v, adv, ret, generator=torch.Generator().manual_seed(1))
        with torch.no_grad():
            v2, logp2, _ = L.policy.evaluate(obs, act)
        outs.append(L.flat_parameters())
        # probability of positively-advantaged actions went up on average
        assert ((logp2 - logp) * adv).mean() > 0
        assert torch.nn.functional.mse_loss(v2, ret) < torch.nn.functional.mse_loss(v, ret)
    assert torch.equal(outs[0], outs[1])


def test_target_kl_early_stop():
    cfg = PPOConfig(batch_size=64, n_epochs=50, target_kl=1e-4, learning_rate=1e-2)
    L = PPOLearner(13, 4, cfg)
    obs, act, adv, ret = _synthetic(0)
    with torch.no_grad():
        v, logp, _ = L.policy.evaluate(obs, act)
    out = L.update(obs, act, logp, v, adv, ret, generator=torch.Generator().manual_seed(1))
    assert out["early_stop"] and out["epochs"] < 50


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = PPOConfig(batch_size=256, n_epochs=1, target_kl=None, max_grad_norm=1e9)
        L = PPOLearner(13, 4, cfg)
        p0 = L.flat_parameters().clone()
        obs, act, adv, ret = _synthetic(rank)
        with torch.no_grad():
            v, logp, _ = L.policy.evaluate(obs, act)
        out = L.update(obs, act, logp, v, adv, ret, generator=torch.Generator().manual_seed(5))   # one minibatch = whole shard
        torch.save({"p0": p0, "p1": L.flat_parameters(), "calls": L.allreduce_calls, "out": out}, os.path.join(tmp, f"r{rank}.pt"))
        # synchronised early stop: only rank 1 exceeds the KL threshold, both must stop
        cfg2 = PPOConfig(batch_size=64, n_epochs=30, target_kl=(1e-5 if rank == 1 else 1e9), learning_rate=1e-2)
        L2 = PPOLearner(13, 4, cfg2)
        with torch.no_grad():
            v, logp, _ = L2.policy.evaluate(obs, act)
        out2 = L2.update(obs, act, logp, v, adv, ret, generator=torch.Generator().manual_seed(5))
        torch.save({"out2": out2, "p": L2.flat_parameters(), "calls": L2.allreduce_calls}, os.path.join(tmp, f"s{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_gradient_allreduce(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(os.path.join(tmp_path, f"r{k}.pt"), weights_only=False) for k in range(world)]
    assert torch.equal(r[0]["p0"], r[1]["p0"])                      # same seed -> identical initial parameters
    assert torch.equal(r[0]["p1"], r[1]["p1"])                      # identical after the all-reduced step
    assert r[0]["calls"] == r[1]["calls"] == 1                      # ONE bucket per optimiser step
    # the step equals Adam on the mean of the two ranks' gradients
    cfg = PPOConfig(batch_size=256, n_epochs=1, target_kl=None, max_grad_norm=1e9)
    L = PPOLearner(13, 4, cfg)
    grads = []
    for k in range(world):
        obs, act, adv, ret = _synthetic(k)
        with torch.no_grad():
            v, logp, _ = L.policy.evaluate(obs, act)
        vals, lp, ent = L.policy.evaluate(obs, act)
        a = (adv - adv.mean()) / (adv.std() + 1e-8)
        ratio = torch.exp(lp - logp)
        pg = -torch.min(a * ratio, a * ratio.clamp(0.8, 1.2)).mean()
        vp = v + (vals - v).clamp(-0.3, 0.3)
        loss = pg + 0.02 * (-ent.mean()) + 0.5 * torch.nn.functional.mse_loss(ret, vp)
        L.opt.zero_grad()
        loss.backward()
        grads.append(torch.cat([p.grad.reshape(-1).clone() for p in L.params]))
    mean_grad = (grads[0] + grads[1]) / 2
    off = 0
    for p in L.params:
        p.grad.copy_(mean_grad[off:off + p.numel()].view_as(p))
        off += p.numel()
    L.opt.step()
    torch.testing.assert_close(L.flat_parameters(), r[0]["p1"], rtol=1e-5, atol=1e-7)
    s = [torch.load(os.path.join(tmp_path, f"s{k}.pt"), weights_only=False) for k in range(world)]
    assert s[0]["out2"]["early_stop"] and s[1]["out2"]["early_stop"]
    assert s[0]["out2"]["minibatches"] == s[1]["out2"]["minibatches"]
    # the early-stop vote rides in the gradient bucket: ONE collective per minibatch, none besides
    assert s[0]["calls"] == s[1]["calls"] == s[0]["out2"]["minibatches"]
    assert torch.equal(s[0]["p"], s[1]["p"])
    assert torch.equal(s[0]["p"], s[1]["p"])
