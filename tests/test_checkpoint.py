"""SB3-format checkpoint I/O (no GPU): archives written here have SB3's member and key names; an archive built the
way SB3 builds one (ActorCriticPolicy.state_dict() key names, parameter order log_std -> mlp_extractor -> heads)
loads into the torch-native policy and produces the same actions / values."""
import io
import json
import zipfile

import pytest
import torch

from drl_dronenavigation_b200.checkpoint import load_sb3_zip, policy_to_sb3_state_dict, save_sb3_zip
from drl_dronenavigation_b200.ppo import PPOConfig, PPOLearner

SB3_KEYS = ["log_std"] + [f"mlp_extractor.{net}.{i}.{leaf}" for net in ("policy_net", "value_net") for i in (0, 2, 4)
                          for leaf in ("weight", "bias")] + ["action_net.weight", "action_net.bias", "value_net.weight", "value_net.bias"]


def _sb3_like_archive(path, seed=7):
    """What `PPO("MlpPolicy", env, policy_kwargs=dict(activation_fn=Tanh, net_arch=dict(pi=[512,512,256], vf=[512,512,256]))).save()`
    puts into policy.pth, with random weights."""
    g = torch.Generator().manual_seed(seed)
    shapes = {"log_std": (4,), "action_net.weight": (4, 256), "action_net.bias": (4,), "value_net.weight": (1, 256), "value_net.bias": (1,)}
    for net in ("policy_net", "value_net"):
        for i, (o, n) in zip((0, 2, 4), ((512, 13), (512, 512), (256, 512))):
            shapes[f"mlp_extractor.{net}.{i}.weight"] = (o, n)
            shapes[f"mlp_extractor.{net}.{i}.bias"] = (o,)
    sd = {k: torch.randn(*shapes[k], generator=g) * 0.1 for k in SB3_KEYS}
    with zipfile.ZipFile(path, "w") as zf:
        buf = io.BytesIO(); torch.save(sd, buf); zf.writestr("policy.pth", buf.getvalue())
        zf.writestr("data", json.dumps({"policy_class": {":type:": "<class 'abc.ABCMeta'>", ":serialized:": "gASV..."}, "n_steps": 4096}))
        zf.writestr("_stable_baselines3_version", "2.3.2")
    return sd


def test_written_archive_has_sb3_members_and_keys(tmp_path):
    L = PPOLearner(13, 4, PPOConfig())
    p = save_sb3_zip(str(tmp_path / "best_model"), L)
    assert p.endswith("best_model.zip")
    with zipfile.ZipFile(p) as zf:
        assert {"data", "policy.pth", "policy.optimizer.pth", "pytorch_variables.pth", "_stable_baselines3_version",
                "system_info.txt"} <= set(zf.namelist())
        sd = torch.load(io.BytesIO(zf.read("policy.pth")), weights_only=True)
        data = json.loads(zf.read("data"))
    assert sorted(sd) == sorted(SB3_KEYS)
    assert sd["mlp_extractor.policy_net.0.weight"].shape == (512, 13) and sd["value_net.weight"].shape == (1, 256)
    assert data["policy_kwargs"]["net_arch"] == {"pi": [512, 512, 256], "vf": [512, 512, 256]} and data["n_epochs"] == 10


def test_reference_style_archive_loads_and_reproduces_outputs(tmp_path):
    path = str(tmp_path / "best_model.zip")
    sd = _sb3_like_archive(path)
    L = PPOLearner(13, 4, PPOConfig())
    load_sb3_zip(path, L)
    obs = torch.randn(9, 13)
    # the SB3 forward pass written out on the raw tensors: latent = tanh MLP, mean = action_net(latent_pi), value = value_net(latent_vf)
    def mlp(x, net):
        for i in (0, 2, 4):
            x = torch.tanh(x @ sd[f"mlp_extractor.{net}.{i}.weight"].T + sd[f"mlp_extractor.{net}.{i}.bias"])
        return x
    mean = mlp(obs, "policy_net") @ sd["action_net.weight"].T + sd["action_net.bias"]
    value = (mlp(obs, "value_net") @ sd["value_net.weight"].T + sd["value_net.bias"]).squeeze(-1)
    a, _, v = L.policy.act(obs, deterministic=True)
    torch.testing.assert_close(a, mean, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(v, value, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(L.policy.log_std.detach(), sd["log_std"])


def test_round_trip_with_optimizer_state(tmp_path):
    cfg = PPOConfig(batch_size=64, n_epochs=1, target_kl=None)
    A = PPOLearner(13, 4, cfg)
    g = torch.Generator().manual_seed(0)
    obs, act, adv, ret = torch.randn(128, 13, generator=g), torch.rand(128, 4, generator=g), torch.randn(128, generator=g), torch.randn(128, generator=g)
    with torch.no_grad():
        v, logp, _ = A.policy.evaluate(obs, act)
    A.update(obs, act, logp, v, adv, ret, generator=torch.Generator().manual_seed(1))
    p = save_sb3_zip(str(tmp_path / "success_model.zip"), A)
    B = PPOLearner(13, 4, PPOConfig(batch_size=64, n_epochs=1, target_kl=None, seed=99))
    assert not torch.equal(A.flat_parameters(), B.flat_parameters())
    load_sb3_zip(p, B, load_optimizer=True)
    assert torch.equal(A.flat_parameters(), B.flat_parameters())
    # same Adam moments -> the next update lands on the same parameters
    for L in (A, B):
        L.update(obs, act, logp, v, adv, ret, generator=torch.Generator().manual_seed(2))
    torch.testing.assert_close(A.flat_parameters(), B.flat_parameters(), rtol=1e-6, atol=1e-8)


def test_wrong_architecture_is_rejected(tmp_path):
    path = str(tmp_path / "m.zip")
    _sb3_like_archive(path)
    L = PPOLearner(13, 4, PPOConfig(pi_arch=(64, 64, 64), vf_arch=(64, 64, 64)))
    with pytest.raises(ValueError, match="shape"):
        load_sb3_zip(path, L)
    assert sorted(policy_to_sb3_state_dict(L.policy)) == sorted(SB3_KEYS)


# ---- SAC archives and the replay buffer (PBDroneSimulator.py:355-362,704-710; Sol/Utilities/Callbacks.py:13-39) -----------------
SAC_SB3_KEYS = ([f"actor.latent_pi.{i}.{leaf}" for i in (0, 2) for leaf in ("weight", "bias")]
                + [f"actor.{h}.{leaf}" for h in ("mu", "log_std") for leaf in ("weight", "bias")]
                + [f"{c}.qf{q}.{i}.{leaf}" for c in ("critic", "critic_target") for q in (0, 1) for i in (0, 2, 4, 6) for leaf in ("weight", "bias")])


def test_sac_archive_has_sb3_members_and_keys_and_round_trips(tmp_path):
    from drl_dronenavigation_b200.sac import SACConfig, SACLearner
    cfg = SACConfig(cuda_graph=False)
    A = SACLearner(13, 4, cfg)
    obs, act = torch.randn(32, 13), torch.rand(32, 4) * 2 - 1
    for _ in range(2):      # two updates so the optimisers and log_ent_coef carry state
        A.update((obs, act, torch.randn(32), torch.randn(32, 13), torch.zeros(32)))
    p = save_sb3_zip(str(tmp_path / "success_model"), A)
    with zipfile.ZipFile(p) as zf:
        assert {"data", "policy.pth", "actor.optimizer.pth", "critic.optimizer.pth", "ent_coef_optimizer.pth", "pytorch_variables.pth",
                "_stable_baselines3_version"} <= set(zf.namelist())
        sd = torch.load(io.BytesIO(zf.read("policy.pth")), weights_only=True)
        data = json.loads(zf.read("data"))
    assert sorted(sd) == sorted(SAC_SB3_KEYS)
    assert sd["actor.latent_pi.0.weight"].shape == (256, 13) and sd["critic.qf1.0.weight"].shape == (256, 17) and sd["critic.qf0.6.weight"].shape == (1, 128)
    assert data["policy_kwargs"]["net_arch"] == {"pi": [256, 256], "qf": [256, 256, 128]} and data["algo"] == "SAC"
    B = SACLearner(13, 4, SACConfig(cuda_graph=False, seed=99))
    load_sb3_zip(p, B, load_optimizer=True)
    with torch.no_grad():
        assert torch.equal(A.actor(obs, deterministic=True)[0], B.actor(obs, deterministic=True)[0])
        for qa, qb in zip(A.critic(obs, act), B.critic(obs, act)):
            assert torch.equal(qa, qb)
        for qa, qb in zip(A.critic_target(obs, act), B.critic_target(obs, act)):
            assert torch.equal(qa, qb)
    assert torch.equal(A.log_ent_coef, B.log_ent_coef)
    # same optimiser state -> the next update moves both learners identically
    batch = (obs, act, torch.randn(32), torch.randn(32, 13), torch.zeros(32))
    ga, gb = torch.Generator().manual_seed(1), torch.Generator().manual_seed(1)
    A.update(batch, generator=ga); B.update(batch, generator=gb)
    assert torch.allclose(A.flat_parameters(), B.flat_parameters(), atol=1e-7)


def test_sb3_style_sac_archive_loads(tmp_path):
    from drl_dronenavigation_b200.sac import SACConfig, SACLearner
    L = SACLearner(13, 4, SACConfig(cuda_graph=False))
    g = torch.Generator().manual_seed(3)
    own = {k: v for pre, mod in (("actor", L.actor), ("critic", L.critic), ("critic_target", L.critic_target)) for k, v in
           ((f"{pre}.{k}", v) for k, v in mod.state_dict().items())}
    from drl_dronenavigation_b200.checkpoint import _sac_key
    sd = {_sac_key(k): torch.randn(v.shape, generator=g) * 0.1 for k, v in own.items()}
    assert sorted(sd) == sorted(SAC_SB3_KEYS)
    path = str(tmp_path / "best_model.zip")
    with zipfile.ZipFile(path, "w") as zf:
        buf = io.BytesIO(); torch.save(sd, buf); zf.writestr("policy.pth", buf.getvalue())
        buf = io.BytesIO(); torch.save({"log_ent_coef": torch.tensor([-1.25])}, buf); zf.writestr("pytorch_variables.pth", buf.getvalue())
    load_sb3_zip(path, L)
    obs = torch.randn(5, 13)
    h = torch.relu(torch.relu(obs @ sd["actor.latent_pi.0.weight"].T + sd["actor.latent_pi.0.bias"]) @ sd["actor.latent_pi.2.weight"].T
                   + sd["actor.latent_pi.2.bias"])
    want = torch.tanh(h @ sd["actor.mu.weight"].T + sd["actor.mu.bias"])
    with torch.no_grad():
        assert torch.allclose(L.actor(obs, deterministic=True)[0], want, atol=1e-6)
    assert float(L.log_ent_coef) == pytest.approx(-1.25)
    bad = dict(sd); bad["actor.mu.weight"] = torch.zeros(4, 128)
    with zipfile.ZipFile(path, "w") as zf:
        buf = io.BytesIO(); torch.save(bad, buf); zf.writestr("policy.pth", buf.getvalue())
    with pytest.raises(ValueError, match="net_arch"):
        load_sb3_zip(path, L)


def test_replay_buffer_save_and_load(tmp_path):
    from drl_dronenavigation_b200.sac import ReplayBuffer
    N, D = 4, 13
    A = ReplayBuffer(6 * N, N, D, 4, "cpu")           # 6 steps of capacity
    g = torch.Generator().manual_seed(0)
    rows = []
    for t in range(9):                                # wraps around: steps 3..8 survive
        o, o2, a = torch.randn(N, D, generator=g), torch.randn(N, D, generator=g), torch.rand(N, 4, generator=g)
        r = torch.randn(N, generator=g)
        done = torch.tensor([0, 1, 2, 3], dtype=torch.uint8) if t % 4 == 0 else torch.zeros(N, dtype=torch.uint8)
        term = torch.randn(N, D, generator=g)
        A.add(o, o2, a, r, done, term)
        rows.append((o, torch.where((done != 0).unsqueeze(-1), term, o2), a, r, ((done & 1) != 0).float()))
    path = A.save(str(tmp_path / "replay_buffer.pkl"))
    import pickle
    d = pickle.load(open(path, "rb"))
    assert d["observations"].shape == (6, N, D) and d["full"] and d["pos"] == 3 and d["n_envs"] == N
    B = ReplayBuffer(6 * N, N, D, 4, "cpu")
    assert B.load(path) == 6 * N and len(B) == 6 * N and B.full
    for k, (o, o2, a, r, dn) in enumerate(rows[3:]):   # oldest -> newest after the reload
        assert torch.equal(B.obs[k], o) and torch.equal(B.next_obs[k], o2) and torch.equal(B.act[k], a)
        assert torch.equal(B.rew[k], r) and torch.equal(B.done[k], dn)
    C = ReplayBuffer(4 * N, N, D, 4, "cpu")            # smaller buffer keeps the most recent steps
    assert C.load(path) == 4 * N and torch.equal(C.obs[0], rows[5][0]) and torch.equal(C.obs[3], rows[8][0])
    P = ReplayBuffer(8 * N, N, D, 4, "cpu")            # partially filled buffers round-trip too
    P.add(*[rows[0][0], rows[0][1], rows[0][2], rows[0][3]], torch.zeros(N, dtype=torch.uint8), rows[0][1])
    P.save(path)
    Q = ReplayBuffer(8 * N, N, D, 4, "cpu")
    assert Q.load(path) == N and Q.pos == 1 and not Q.full and torch.equal(Q.obs[0], rows[0][0])
    with pytest.raises(ValueError, match="n_envs"):
        ReplayBuffer(8, 2, D, 4, "cpu").load(path)


def test_replay_buffer_load_honours_sb3_timeouts(tmp_path):
    """An SB3 ReplayBuffer pickle stores `dones` (terminated OR truncated) next to `timeouts`; SB3 samples dones * (1 - timeouts).
    Loading such an object must not turn time-limit truncations into terminal states (no bootstrap)."""
    import pickle
    import numpy as np
    from drl_dronenavigation_b200.sac import ReplayBuffer
    N, D, T = 2, 13, 3
    rng = np.random.default_rng(0)
    obj = {"observations": rng.normal(size=(T, N, D)).astype(np.float32), "next_observations": rng.normal(size=(T, N, D)).astype(np.float32),
           "actions": rng.uniform(-1, 1, size=(T, N, 4)).astype(np.float32), "rewards": rng.normal(size=(T, N)).astype(np.float32),
           "dones": np.array([[1, 0], [1, 1], [0, 0]], np.float32), "timeouts": np.array([[1, 0], [0, 1], [0, 0]], np.float32),
           "pos": 3, "full": False, "buffer_size": 8, "n_envs": N}
    path = str(tmp_path / "sb3_like.pkl")
    pickle.dump(obj, open(path, "wb"))
    B = ReplayBuffer(8 * N, N, D, 4, "cpu")
    assert B.load(path) == T * N
    assert B.done[:T].tolist() == [[0.0, 0.0], [1.0, 0.0], [0.0, 0.0]]
