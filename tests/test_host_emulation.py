"""CPU-side check of the DEVICE step logic: csrc/dn_device.cuh (the source the CUDA kernel is built
from) is compiled for the host by g++ (tests/host_emu) and run in lock-step against the oracle with
the same tolerances as the GPU parity tests.  This is a test tool -- it catches logic regressions
in the kernel source on machines without a GPU; the real parity tests are the -m gpu ones."""
import numpy as np
import pytest

from tests import parity_utils as PU

HOVER = 0.092227


def _make(track, N, S, max_steps=4096, physics=0, oracle_physics="dyn"):
    from oracle.dyn_oracle import OracleWorker, make_reference_env
    from tests.host_emu import HostEmuEnv
    ctrl = 240 // S
    mk = lambda: make_reference_env(track, pyb_freq=240, ctrl_freq=ctrl, max_steps=max_steps, physics=oracle_physics)
    ref = mk()
    env = HostEmuEnv(N, ref._target_points, max_steps=max_steps, aviary_dim=ref._aviary_dim, initial_xyzs=ref.INIT_XYZS,
                     pyb_freq=240, ctrl_freq=ctrl, circle=(track == "circle"), include_distance=True,
                     normalize_actions=True, physics=physics)
    workers = [OracleWorker(mk(), normalize_obs=False) for _ in range(N)]
    for w in workers:
        w.reset()
    return env, workers


def _actions(mode, T, N, seed):
    u = np.random.default_rng(seed).uniform(-1, 1, size=(T, N, 4))
    a = {"saturating": u, "hover_band": HOVER + 0.002 * u, "mixed": HOVER + 0.006 * u}[mode]
    return a.astype(np.float32)


@pytest.mark.parametrize("track,S,mode,N,T", [
    ("circle", 1, "saturating", 8, 240),
    ("circle", 8, "mixed", 8, 60),
    ("reaching", 8, "saturating", 8, 60),
    ("reaching", 1, "hover_band", 4, 240),
])
def test_device_logic_lockstep(track, S, mode, N, T):
    env, workers = _make(track, N, S)
    rep = PU.run_lockstep(env, workers, _actions(mode, T, N, seed=S * 7 + len(mode)), resync_every=240 // S)
    print(f"\n[emu {track} S={S} {mode}] {rep}")
    assert rep.near_ties <= 2
    env.close()


def test_device_action_map_bit_exact():
    """Constant-divisor division by FMA residual (div_const_rn) == numpy's float32 division, bit for bit,
    over every float32 in the pass-through band of the THRUST map and a random sample outside it."""
    from oracle.dyn_oracle import physical_action_bounds, rescale_action_batch, thrust_to_rpm_batch
    env, _ = _make("circle", 1, 1)
    b = physical_action_bounds()
    band = np.arange(np.float32(0.0895).view(np.int32), np.float32(0.0975).view(np.int32) + 1, dtype=np.int32).view(np.float32)
    rnd = np.random.default_rng(0).uniform(-1.2, 1.2, size=200_000).astype(np.float32)
    a = np.concatenate([band, rnd])
    a = a[: (a.size // 4) * 4]
    ref = thrust_to_rpm_batch(rescale_action_batch(a, b), b).reshape(-1)
    got = env.action_to_rpm(a)
    np.testing.assert_array_equal(got.view(np.int32), ref.view(np.int32))
    env.close()


@pytest.mark.parametrize("physics,oracle_physics", [(1, "dyn_drag"), (2, "dyn_gnd"), (3, "dyn_gnd_drag")])
def test_device_logic_physics_addons(physics, oracle_physics):
    env, workers = _make("circle", 6, 8, physics=physics, oracle_physics=oracle_physics)
    rep = PU.run_lockstep(env, workers, _actions("mixed", 45, 6, seed=physics), resync_every=30)
    print(f"\n[emu physics={physics}] {rep}")
    env.close()
