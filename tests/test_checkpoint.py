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
