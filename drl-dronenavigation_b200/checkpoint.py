"""SB3-format checkpoint I/O for the PPO policy.

The reference saves and reloads Stable-Baselines3 archives -- ``best_model.zip`` / ``success_model.zip`` written by
``EvalCallback`` / ``model.save`` (Sol/Model/PBDroneSimulator.py:719-746) and read back by ``PPO.load`` in
``test_saved`` (:438-572) and ``run_full(cont)`` (:132-134).  An SB3 archive is a zip with

    data                          JSON of the constructor arguments (spaces etc. as base64 cloudpickle)
    policy.pth                    torch state_dict of the ActorCriticPolicy
    policy.optimizer.pth          torch state_dict of its Adam optimiser
    pytorch_variables.pth         extra tensors (none for PPO)
    _stable_baselines3_version    text
    system_info.txt               text

This module reads ``policy.pth`` (and the optimiser, when present) of such an archive into the torch-native
``ActorCritic`` of ``ppo.py`` and writes archives with the same member names and the same ``policy.pth`` key names, so
policies move in both directions:  reference-trained ``*.zip`` -> GPU environment, and a policy trained here ->
``ActorCriticPolicy.load_state_dict(torch.load(zf.open("policy.pth")))`` in an SB3 installation.  The ``data`` member
written here is plain JSON of the hyper-parameters (no cloudpickled spaces: gymnasium is not installable in this
image), which ``PPO.load`` needs ``custom_objects`` for; the tensor members are exact.

Key map for ``net_arch=dict(pi=[512, 512, 256], vf=[512, 512, 256])``, ``share_features_extractor=False``
(PBDroneSimulator.py:251-258; SB3 MlpExtractor / ActorCriticPolicy naming):

    mlp_extractor.policy_net.{0,2,4}.{weight,bias}  <->  pi.{0,2,4}.{weight,bias}
    action_net.{weight,bias}                        <->  pi.6.{weight,bias}
    mlp_extractor.value_net.{0,2,4}.{weight,bias}   <->  vf.{0,2,4}.{weight,bias}
    value_net.{weight,bias}                         <->  vf.6.{weight,bias}
    log_std                                         <->  log_std
"""
from __future__ import annotations

import io
import json
import zipfile
from typing import Dict, Optional

import torch

SB3_VERSION_TAG = "2.3.2"        # written to _stable_baselines3_version (format unchanged since SB3 1.x)


def _key_map(n_hidden_pi: int, n_hidden_vf: int) -> Dict[str, str]:
    """ours -> SB3."""
    m = {"log_std": "log_std"}
    for ours, theirs, head, n in (("pi", "policy_net", "action_net", n_hidden_pi), ("vf", "value_net", "value_net", n_hidden_vf)):
        for layer in range(n):
            for leaf in ("weight", "bias"):
                m[f"{ours}.{2 * layer}.{leaf}"] = f"mlp_extractor.{theirs}.{2 * layer}.{leaf}"
        for leaf in ("weight", "bias"):
            m[f"{ours}.{2 * n}.{leaf}"] = f"{head}.{leaf}"
    return m


def policy_to_sb3_state_dict(policy) -> Dict[str, torch.Tensor]:
    sd = policy.state_dict()
    n_pi = sum(1 for k in sd if k.startswith("pi.") and k.endswith(".weight")) - 1
    n_vf = sum(1 for k in sd if k.startswith("vf.") and k.endswith(".weight")) - 1
    km = _key_map(n_pi, n_vf)
    return {km[k]: v.detach().cpu().clone() for k, v in sd.items()}


def sb3_state_dict_to_policy(policy, sb3_sd: Dict[str, torch.Tensor], strict: bool = True) -> None:
    sd = policy.state_dict()
    n_pi = sum(1 for k in sd if k.startswith("pi.") and k.endswith(".weight")) - 1
    n_vf = sum(1 for k in sd if k.startswith("vf.") and k.endswith(".weight")) - 1
    km = _key_map(n_pi, n_vf)
    new = {}
    for ours, theirs in km.items():
        if theirs not in sb3_sd:
            if strict:
                raise KeyError(f"{theirs} missing from the SB3 policy state_dict (has: {sorted(sb3_sd)[:6]} ...)")
            continue
        t = sb3_sd[theirs]
        if tuple(t.shape) != tuple(sd[ours].shape):
            raise ValueError(f"{theirs}: shape {tuple(t.shape)} does not match this policy's {tuple(sd[ours].shape)} "
                             "(net_arch differs from PBDroneSimulator.py:251-258?)")
        new[ours] = t.to(sd[ours].device, sd[ours].dtype)
    extra = [k for k in sb3_sd if k not in km.values() and not k.startswith(("features_extractor", "pi_features_extractor",
                                                                              "vf_features_extractor"))]
    if strict and extra:
        raise KeyError(f"unexpected keys in the SB3 policy state_dict: {extra[:6]}")
    policy.load_state_dict(new, strict=strict)


def _save_tensor_member(zf: zipfile.ZipFile, name: str, obj) -> None:
    buf = io.BytesIO()
    torch.save(obj, buf)
    zf.writestr(name, buf.getvalue())


def save_sb3_zip(path: str, learner, extra: Optional[dict] = None) -> str:
    """Writes an SB3-layout archive of a ``PPOLearner`` (policy + Adam state + hyper-parameters)."""
    cfg = learner.cfg
    sd = policy_to_sb3_state_dict(learner.policy)
    km = _key_map(len(cfg.pi_arch), len(cfg.vf_arch))
    # SB3's optimiser state_dict indexes parameters by position in policy.parameters(); keep ours and record the names
    opt_sd = learner.opt.state_dict()
    opt_state = {"state": {i: {k: (v.detach().cpu() if torch.is_tensor(v) else v) for k, v in st.items()}
                           for i, st in opt_sd["state"].items()},
                 "param_groups": opt_sd["param_groups"],
                 "param_names": [km[n] for n, _ in learner.policy.named_parameters()]}
    data = {"policy_class": "stable_baselines3.common.policies.ActorCriticPolicy", "algo": "PPO",
            "n_steps": cfg.n_steps, "batch_size": cfg.batch_size, "n_epochs": cfg.n_epochs, "gamma": cfg.gamma,
            "gae_lambda": cfg.gae_lambda, "ent_coef": cfg.ent_coef, "vf_coef": cfg.vf_coef, "clip_range": cfg.clip_range,
            "clip_range_vf": cfg.clip_range_vf, "normalize_advantage": cfg.normalize_advantage,
            "max_grad_norm": cfg.max_grad_norm, "target_kl": cfg.target_kl, "learning_rate": cfg.learning_rate,
            "policy_kwargs": {"activation_fn": "torch.nn.Tanh", "net_arch": {"pi": list(cfg.pi_arch), "vf": list(cfg.vf_arch)},
                              "share_features_extractor": False, "log_std_init": cfg.log_std_init},
            "observation_space": {"type": "Box", "shape": [learner.policy.pi[0].in_features], "dtype": "float32"},
            "action_space": {"type": "Box", "low": -1.0, "high": 1.0, "shape": [learner.policy.log_std.numel()], "dtype": "float32"},
            "n_updates": learner.n_updates, "seed": cfg.seed, "written_by": "drl_dronenavigation_b200"}
    if extra:
        data.update(extra)
    if not path.endswith(".zip"):
        path += ".zip"
    with zipfile.ZipFile(path, "w", zipfile.ZIP_DEFLATED) as zf:
        zf.writestr("data", json.dumps(data, indent=2))
        _save_tensor_member(zf, "policy.pth", sd)
        _save_tensor_member(zf, "policy.optimizer.pth", opt_state)
        _save_tensor_member(zf, "pytorch_variables.pth", None)
        zf.writestr("_stable_baselines3_version", SB3_VERSION_TAG)
        zf.writestr("system_info.txt", f"torch {torch.__version__}\n")
    return path


def load_sb3_zip(path: str, learner, load_optimizer: bool = False, strict: bool = True) -> dict:
    """Loads ``policy.pth`` of an SB3 archive (the reference's ``best_model.zip``) into ``learner.policy``.
    Returns the parsed ``data`` member (or {} when it is not plain JSON-decodable)."""
    with zipfile.ZipFile(path) as zf:
        names = set(zf.namelist())
        if "policy.pth" not in names:
            raise KeyError(f"{path}: no policy.pth member (members: {sorted(names)})")
        sb3_sd = torch.load(io.BytesIO(zf.read("policy.pth")), map_location="cpu", weights_only=True)
        sb3_state_dict_to_policy(learner.policy, sb3_sd, strict=strict)
        if load_optimizer and "policy.optimizer.pth" in names:
            opt = torch.load(io.BytesIO(zf.read("policy.optimizer.pth")), map_location="cpu", weights_only=False)
            own = learner.opt.state_dict()
            if len(opt.get("state", {})) == len(list(learner.policy.parameters())):
                # SB3 orders parameters as policy.parameters() does: log_std, policy_net, value_net, action_net, value_net head;
                # ours: pi..., vf..., log_std -- re-index by name when the names were recorded, else by shape-compatible order
                names_theirs = opt.get("param_names")
                if names_theirs is not None:
                    km = _key_map(len(learner.cfg.pi_arch), len(learner.cfg.vf_arch))
                    order = {km[n]: i for i, (n, _) in enumerate(learner.policy.named_parameters())}
                    state = {order[nm]: opt["state"][j] for j, nm in enumerate(names_theirs) if j in opt["state"]}
                    own["state"] = state
                    learner.opt.load_state_dict(own)
                    # load_state_dict re-creates the state tensors: captured CUDA graphs (if any) must be rebuilt
                    if hasattr(learner, "_graphs"):
                        learner._graphs = None
        data = {}
        if "data" in names:
            try:
                data = json.loads(zf.read("data").decode())
            except Exception:  # noqa: BLE001
                data = {}
    return data
