"""Host-side logic of the SAC learner (no GPU): network shapes of the reference's SAC branch, the replay buffer's
terminal-observation / time-limit handling, one update step against a hand-written restatement of SB3's SAC.train,
and the N>1 path (two flat-gradient all-reduces per gradient step) on the gloo backend."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from drl_dronenavigation_b200.sac import Actor, Critics, ReplayBuffer, SACConfig, SACLearner


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_networks_match_reference_sac_branch():
    cfg = SACConfig()      # PBDroneSimulator.py:297-327
    L = SACLearner(13, 4, cfg)
    n_actor = sum(p.numel() for p in L.actor.parameters())
    n_critic = sum(p.numel() for p in L.critic.parameters())
    assert n_actor == (13 * 256 + 256) + (256 * 256 + 256) + 2 * (256 * 4 + 4)
    assert n_critic == 2 * ((17 * 256 + 256) + (256 * 256 + 256) + (256 * 128 + 128) + (128 + 1))
    assert L.target_entropy == -4.0 and float(L.log_ent_coef) == 0.0
    assert (cfg.batch_size, cfg.buffer_size, cfg.train_freq, cfg.gradient_steps, cfg.learning_starts) == (1024, 1_048_576, 3, 5, 8192)
    a, logp = L.actor(torch.randn(5, 13))
    assert a.shape == (5, 4) and logp.shape == (5,) and (a.abs() <= 1).all()
    a_det, _ = L.actor(torch.zeros(1, 13), deterministic=True)
    assert torch.allclose(a_det, torch.tanh(L.actor.mu(L.actor.latent(torch.zeros(1, 13)))))


def test_replay_buffer_terminal_obs_and_timeouts():
    N, D = 4, 13
    buf = ReplayBuffer(3 * N, N, D, 4, "cpu")
    assert buf.cap == 3
    obs, nxt, term = torch.zeros(N, D), torch.ones(N, D), torch.full((N, D), 7.0)
    bits = torch.tensor([0, 1, 2, 3], dtype=torch.uint8)      # running, terminated, truncated, both
    buf.add(obs, nxt, torch.zeros(N, 4), torch.arange(N, dtype=torch.float32), bits, term)
    assert len(buf) == N
    # successor of a finished episode is its TERMINAL observation, not the reset observation the VecEnv returned
    np.testing.assert_array_equal(buf.next_obs[0, :, 0].numpy(), [1, 7, 7, 7])
    # time-limit truncation alone is not a termination (the target bootstraps through it)
    np.testing.assert_array_equal(buf.done[0].numpy(), [0, 1, 0, 1])
    for _ in range(3):
        buf.add(obs, nxt, torch.zeros(N, 4), torch.zeros(N), torch.zeros(N, dtype=torch.uint8), term)
    assert buf.full and len(buf) == 3 * N and buf.pos == 1
    o, a, r, no, d = buf.sample(32, generator=torch.Generator().manual_seed(0))
    assert o.shape == (32, D) and a.shape == (32, 4) and r.shape == (32,) and no.shape == (32, D) and d.shape == (32,)


def _batch(rank, B=128):
    g = torch.Generator().manual_seed(100 + rank)
    return (torch.randn(B, 13, generator=g), torch.rand(B, 4, generator=g) * 2 - 1, torch.randn(B, generator=g),
            torch.randn(B, 13, generator=g), (torch.rand(B, generator=g) < 0.1).float())


def _reference_step(L: SACLearner, batch, seed):
    """SB3 SAC.train for one gradient step, written out with plain autograd / separate optimisers."""
    cfg = L.cfg
    obs, act, rew, next_obs, done = batch
    gen = torch.Generator().manual_seed(seed)
    actions_pi, log_prob = L.actor(obs, generator=gen)
    ent_coef = torch.exp(L.log_ent_coef.detach())
    ent_coef_loss = -(L.log_ent_coef * (log_prob + L.target_entropy).detach()).mean()
    with torch.no_grad():
        next_actions, next_log_prob = L.actor(next_obs, generator=gen)
        next_q = torch.cat([q.unsqueeze(1) for q in L.critic_target(next_obs, next_actions)], dim=1).min(dim=1).values
        target_q = rew + (1 - done) * cfg.gamma * (next_q - ent_coef * next_log_prob)
    grads = {}
    g = torch.autograd.grad(ent_coef_loss, [L.log_ent_coef])
    grads["ent"] = g[0]
    current_q = L.critic(obs, act)
    critic_loss = 0.5 * sum(torch.nn.functional.mse_loss(q, target_q) for q in current_q)
    grads["critic"] = torch.autograd.grad(critic_loss, list(L.critic.parameters()))
    return grads, float(critic_loss), float(ent_coef_loss)


def test_update_matches_sb3_restatement_and_learns():
    cfg = SACConfig(batch_size=128)
    L, R = SACLearner(13, 4, cfg), SACLearner(13, 4, cfg)
    batch = _batch(0)
    grads, closs, eloss = _reference_step(R, batch, seed=3)
    out = L.update(batch, generator=torch.Generator().manual_seed(3))
    assert abs(float(out["critic_loss"]) - closs) < 1e-6 and abs(float(out["ent_coef_loss"]) - eloss) < 1e-6
    # the critic / log-alpha gradients that were applied are the restatement's
    off = 0
    flat_ref = torch.cat([g.reshape(-1) for g in grads["critic"]] + [grads["ent"].reshape(-1)])
    torch.testing.assert_close(L.g_critic.flat, flat_ref, rtol=1e-5, atol=1e-7)
    # targets moved by tau towards the (updated) critics
    for tp, p, rp in zip(L.critic_target.parameters(), L.critic.parameters(), R.critic_target.parameters()):
        torch.testing.assert_close(tp, (1 - cfg.tau) * rp + cfg.tau * p, rtol=1e-5, atol=1e-7)
    # repeated updates on one batch drive the critic loss down and alpha below its initial 1
    first = float(out["critic_loss"])
    for k in range(60):
        out = L.update(batch, generator=torch.Generator().manual_seed(10 + k))
    assert float(out["critic_loss"]) < first and float(out["ent_coef"]) < 1.0


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L = SACLearner(13, 4, SACConfig(batch_size=128))
        p0 = L.flat_parameters().clone()
        for k in range(3):
            L.update(_batch(rank), generator=torch.Generator().manual_seed(7 + rank + k))
        torch.save({"p0": p0, "p1": L.flat_parameters(), "calls": (L.g_critic.calls, L.g_actor.calls),
                    "tgt": torch.cat([p.reshape(-1) for p in L.critic_target.parameters()])}, os.path.join(tmp, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_gradient_allreduce(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(os.path.join(tmp_path, f"r{k}.pt"), weights_only=False) for k in range(world)]
    assert torch.equal(r[0]["p0"], r[1]["p0"])            # same seed -> identical initial parameters
    assert torch.equal(r[0]["p1"], r[1]["p1"])            # different data and noise per rank, identical parameters after
    assert torch.equal(r[0]["tgt"], r[1]["tgt"])
    assert not torch.equal(r[0]["p0"], r[0]["p1"])
    assert r[0]["calls"] == r[1]["calls"] == (3, 3)       # two buckets per gradient step
