"""PBDroneSimulator (the reference's experiment manager, Sol/Model/PBDroneSimulator.py) driven end to end on the CPU: the
device step logic comes from its host build (tests/emu_torch_env.py), the learners run on torch CPU tensors.  Covers the
paths the reference's command line reaches: --agent PPO / SAC, --run_type full / cont, --savemodel (SB3-layout archives,
SAC replay buffer), --tensorboard."""
import argparse
import glob
import os
import zipfile

import pytest
import torch

from drl_dronenavigation_b200 import waypoints as W
from drl_dronenavigation_b200.simulator import PBDroneSimulator, load_most_recent_replay_buffer
from tests.emu_torch_env import EmuTorchEnv


def _args(**kw):
    a = dict(agent="PPO", run_type="full", num_envs=32, max_env_steps=200, total_timesteps=32 * 16 * 3, savemodel=True, pyb_freq=240,
             ctrl_freq=30, rollout_steps=16, minibatch=128, n_epochs=2, clip_rew=False, norm_rew=False, reward_id=0, tensorboard=None,
             model_path=None, max_seconds=None, rank=0, world=1)
    a.update(kw)
    return argparse.Namespace(**a)


@pytest.fixture
def manager(monkeypatch, tmp_path):
    monkeypatch.chdir(tmp_path)

    def make(**kw):
        track = W.Track(W.circle(radius=1, num_points=6, height=1), circle=True)
        sim = PBDroneSimulator(_args(**kw), track)

        def make_device_env(num_envs, normalize_obs=False, device=None, env_id_offset=0):
            return EmuTorchEnv(num_envs, sim.targets, **sim._env_kwargs(None, sim.aviary_dim, True, True))
        sim.make_device_env = make_device_env
        return sim
    return make


def test_ppo_full_run_saves_sb3_archives_and_continues(manager, tmp_path):
    lines = []
    sim = manager(tensorboard=str(tmp_path / "tb"))
    trainer, ev = sim.run_full_training(log=lines.append)
    assert trainer.total_steps >= 32 * 16 * 3 and ev["episodes"] >= 1000 and 0.0 <= ev["success_rate"] <= 1.0
    assert any(l.startswith("final evaluation") for l in lines)
    zips = sorted(glob.glob(str(tmp_path / "Sol" / "model_chkpts" / "PPO_save_*" / "*.zip")))
    assert any(z.endswith("success_model.zip") for z in zips)
    assert glob.glob(str(tmp_path / "tb" / "events.out.tfevents.*"))
    cont = manager(run_type="cont", model_path=[z for z in zips if z.endswith("success_model.zip")][0], savemodel=False,
                   total_timesteps=32 * 16)
    t2, _ = cont.run_full_training(log=lines.append)
    # the continued learner starts from the saved parameters (then trains one iteration on top)
    assert t2.learner.n_updates > 0


def test_sac_full_run_saves_archive_and_replay_buffer_and_continues(manager, tmp_path):
    from drl_dronenavigation_b200 import sac as S
    lines = []
    sim = manager(agent="SAC", total_timesteps=32 * 3 * 12)
    orig = S.SACConfig
    try:
        # the reference's SAC hyper-parameters with a small buffer / early learning start so that updates happen within the test
        S.SACConfig = lambda: orig(learning_starts=32 * 3 * 4, batch_size=64, buffer_size=32 * 64, cuda_graph=False)
        trainer, ev = sim.run_full_training(log=lines.append)
        assert trainer.learner.n_updates > 0 and len(trainer.buffer) > 0 and ev["episodes"] >= 1000
        assert any("critic_loss" in l for l in lines)
        chk = glob.glob(str(tmp_path / "Sol" / "model_chkpts" / "SAC_save_*"))[0]
        with zipfile.ZipFile(os.path.join(chk, "success_model.zip")) as zf:
            assert {"policy.pth", "actor.optimizer.pth", "critic.optimizer.pth", "ent_coef_optimizer.pth", "pytorch_variables.pth"} <= set(zf.namelist())
        assert load_most_recent_replay_buffer(chk) == os.path.join(chk, "replay_buffer.pkl")
        saved = trainer.learner.flat_parameters().clone()
        filled = len(trainer.buffer)
        cont = manager(agent="SAC", run_type="cont", model_path=os.path.join(chk, "success_model.zip"), savemodel=False, total_timesteps=0)
        t2, _ = cont.run_full_training(log=lines.append)
        assert torch.equal(t2.learner.flat_parameters(), saved)          # SAC.load(...) restored actor, critics and log_ent_coef
        assert len(t2.buffer) == filled                                  # model.load_replay_buffer(...)
        assert any(l.startswith("replay buffer") for l in lines)
    finally:
        S.SACConfig = orig


def test_most_recent_replay_buffer_pattern(tmp_path):
    assert load_most_recent_replay_buffer(str(tmp_path)) is None
    (tmp_path / "replay_buffer.pkl").write_bytes(b"")
    assert load_most_recent_replay_buffer(str(tmp_path)).endswith("replay_buffer.pkl")
    for n in (3, 12, 7):
        (tmp_path / f"replay_buffer_{n}.pkl").write_bytes(b"")
    assert load_most_recent_replay_buffer(str(tmp_path)).endswith("replay_buffer_12.pkl")      # PBDroneSimulator.py:998-1017


def test_learning_and_saved_run_types(manager, tmp_path):
    """--run_type learning (PBDroneSimulator.py:574-612) and --run_type saved (:438-572)."""
    sim = manager(num_envs=1, max_env_steps=64, batch_size=32, savemodel=False)
    trainer, out = sim.test_learning(total_timesteps=128, log=lambda *_: None)
    assert trainer.total_steps >= 128 and "approx_kl" in out
    assert [m.out_features for m in trainer.learner.policy.pi if hasattr(m, "out_features")] == [512, 512, 256, 128, 4]
    # a saved PPO archive and a saved SAC archive are both rolled out by test_saved
    from drl_dronenavigation_b200.checkpoint import save_sb3_zip
    from drl_dronenavigation_b200.ppo import PPOConfig, PPOLearner
    from drl_dronenavigation_b200.sac import SACConfig, SACLearner
    p_ppo = save_sb3_zip(str(tmp_path / "ppo_best_model"), PPOLearner(13, 4, PPOConfig()))
    p_sac = save_sb3_zip(str(tmp_path / "sac_best_model"), SACLearner(13, 4, SACConfig(cuda_graph=False)))
    ev = manager(savemodel=False).test_saved(p_ppo, episodes=20)
    assert ev["episodes"] >= 20 and ev["mean_ep_length"] > 0
    ev = manager(agent="SAC", savemodel=False).test_saved(p_sac, episodes=20)
    assert ev["episodes"] >= 20 and ev["mean_found_targets"] >= 0


# ---- N > 1: the manager's data-parallel path (one process per shard, gradients all-reduced) on gloo, world_size 2 --------------
def _dp_worker(rank, world, port, tmp, agent):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    os.chdir(tmp)
    torch.set_num_threads(2)                      # two ranks share the host's cores
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from drl_dronenavigation_b200 import sac as S
        track = W.Track(W.circle(radius=1, num_points=6, height=1), circle=True)
        # rank 1 would stop one iteration later on its own (a later wall-clock cap would do the same): ranks must leave together
        sim = PBDroneSimulator(_args(agent=agent, rank=rank, world=world, savemodel=(rank == 0), num_envs=16,
                                     total_timesteps=(16 * 2 * 16 * 3 if agent == "PPO" else 16 * 2 * 3 * 10)), track)
        offsets = []

        def make_device_env(num_envs, normalize_obs=False, device=None, env_id_offset=0):
            offsets.append(env_id_offset)
            return EmuTorchEnv(num_envs, sim.targets, **sim._env_kwargs(None, sim.aviary_dim, True, True))
        sim.make_device_env = make_device_env
        orig = S.SACConfig
        S.SACConfig = lambda: orig(learning_starts=16 * 2 * 3 * 3, batch_size=64, buffer_size=16 * 64, cuda_graph=False)
        try:
            trainer, ev = sim.run_full_training(log=lambda *_: None)
        finally:
            S.SACConfig = orig
        torch.save({"p": trainer.learner.flat_parameters(), "steps": trainer.total_steps, "offset": offsets[0],
                    "n_updates": trainer.learner.n_updates}, os.path.join(tmp, f"dp_{agent}_{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("agent", ["PPO", "SAC"])
def test_two_rank_gloo_manager_run(agent, tmp_path):
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_dp_worker, args=(2, port, str(tmp_path), agent), nprocs=2, join=True)
    r = [torch.load(os.path.join(tmp_path, f"dp_{agent}_{k}.pt"), weights_only=False) for k in range(2)]
    assert torch.equal(r[0]["p"], r[1]["p"])                          # parameters stay bit-identical across ranks
    assert r[0]["steps"] == r[1]["steps"] and r[0]["n_updates"] == r[1]["n_updates"] > 0
    assert (r[0]["offset"], r[1]["offset"]) == (0, 16)                # contiguous global env-id shards
    assert len(glob.glob(str(tmp_path / "Sol" / "model_chkpts" / f"{agent}_save_*"))) == 1      # one writer per job


def test_profile_flag_wraps_the_training_run(monkeypatch, tmp_path, capsys):
    """--profile t (Sol/Model/simulation_controller.py:111-117): cProfile around run_full_training, cumtime-sorted report."""
    from drl_dronenavigation_b200 import simulation_controller as SC
    monkeypatch.chdir(tmp_path)
    real = PBDroneSimulator.make_device_env

    def make_device_env(self, num_envs, normalize_obs=False, device=None, env_id_offset=0):
        return EmuTorchEnv(num_envs, self.targets, **self._env_kwargs(None, self.aviary_dim, True, True))
    monkeypatch.setattr(PBDroneSimulator, "make_device_env", make_device_env)
    SC.main(["--agent", "PPO", "--run_type", "full", "--num_envs", "16", "--total_timesteps", "512", "--rollout_steps", "16", "--minibatch", "64",
             "--n_epochs", "1", "--max_env_steps", "100", "--savemodel", "f", "--profile", "t", "--ctrl_freq", "30"])
    out = capsys.readouterr().out
    assert "cumulative" in out and "train_iteration" in out and "final evaluation" in out
    assert real is not None


@pytest.mark.parametrize("flags,msg", [(["--lib", "ray"], "out of scope"), (["--wandb", "t"], "tensorboard"), (["--gui", "t"], "GUI"),
                                       (["--capture-video", "t"], "renderer"), (["--vec_normalize", "t"], "VecNormalize"),
                                       (["--vec_check_nan", "t"], "VecCheckNan")])
def test_out_of_scope_flags_fail_loudly(flags, msg):
    from drl_dronenavigation_b200 import simulation_controller as SC
    with pytest.raises(NotImplementedError, match=msg):
        SC.main(flags)
