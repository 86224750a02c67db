"""Mints ``tests/golden/ref_*.npz`` by stepping the REFERENCE'S OWN code (imported unmodified from
``/root/reference``), so that the oracle restatement and the CUDA path are pinned against what the
reference computes rather than against our reading of it.

Run (in the build container only -- /root/reference does not exist on the GPU box):

    python tests/golden/make_ref_golden.py

What executes: ``Sol.Model.Environments.PBDroneEnv.PBDroneEnv`` (``step``, ``reset``, ``rescale_action``,
``_preprocessAction``, ``_computeObs``, ``_clipAndNormalizeState``, ``_computeReward``, ``orientation_reward``,
``smoothness_reward``, ``_computeTerminated``, ``is_out_of_cylinder_bounds``, ``_computeTruncated``,
``_update_state_post_step``), ``Sol.PyBullet.BaseAviary.BaseAviary`` (``step``, ``_dynamics``, ``_integrateQ``,
``_housekeeping``, ``_updateAndStoreKinematicInformation``, ``_parse_urdf_parameters`` on the reference's
own ``cf2x.urdf``), ``Sol.Model.env_utils`` (``cmd2pwm``, ``pwm2rpm``), ``Sol.Model.Environments.normalize``
(``NormalizeObservation``, ``RunningMeanStd``) and ``Sol.Utilities.Waypoints`` (``circle``, ``reaching``, ``Track``).

The three things that are NOT the reference (each unavoidable, each documented in DESIGN.md section 2):

1. third-party packages absent from this image are shimmed (``tests/golden/ref_shims.py``): pybullet becomes a
   state store with Bullet's quaternion maths restated; gymnasium/gym a minimal ``Env``/``Box``;
2. ``BaseAviary.py:418`` overwrites ``self.PHYSICS`` with ``Physics.PYB`` inside the substep loop, which makes
   the DYN branch unreachable.  ``_DynEnv`` below pins the attribute to ``Physics.DYN`` (a read-only property in
   a subclass; the assignment at :418 becomes a no-op), i.e. the loop runs as written minus that line;
3. ``BaseAviary.py:944`` reads the undefined ``self.TIMESTEP``; the subclass defines it as ``PYB_TIMESTEP``.

The SubprocVecEnv worker's auto-reset and the Monitor accumulators are SB3 code (absent); the loop in ``run``
restates that contract (reset on done, returned obs = reset obs, terminal obs kept aside).
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
HOVER = 0.092227

CASES = {
    # name: (track, S, action mode, envs, steps, seed, max_steps, normalize_obs)
    "ref_circle_s1_saturating": ("circle", 1, "saturating", 4, 160, 21, 4096, False),
    "ref_circle_s8_mixed": ("circle", 8, "mixed", 4, 60, 22, 4096, False),
    "ref_circle_s8_saturating": ("circle", 8, "saturating", 4, 40, 26, 4096, False),
    "ref_reaching_s8_saturating": ("reaching", 8, "saturating", 4, 50, 23, 4096, False),
    "ref_reaching_s1_hover": ("reaching", 1, "hover_band", 2, 200, 24, 4096, False),
    "ref_circle_s1_scripted": ("circle", 1, "scripted", 3, 120, 0, 4096, False),
    "ref_circle_s1_truncate": ("circle", 1, "hover", 2, 30, 0, 12, False),
    "ref_circle_s8_normobs": ("circle", 8, "mixed", 3, 60, 25, 4096, True),
    "ref_circle_s1_normobs_resets": ("circle", 1, "saturating", 3, 140, 27, 4096, True),
}


def actions(mode, T, N, seed):
    u = np.random.default_rng(seed).uniform(-1, 1, size=(T, N, 4))
    if mode == "scripted":     # SURVEY 8c scenarios: hover band, max thrust, min thrust
        a = np.zeros((T, N, 4))
        a[:, 0], a[:, 1], a[:, 2] = 0.0922265, 1.0, -1.0
        return a.astype(np.float32)
    if mode == "hover":
        return np.full((T, N, 4), 0.0922265, np.float32)
    return {"saturating": u, "hover_band": HOVER + 0.002 * u, "mixed": HOVER + 0.006 * u}[mode].astype(np.float32)


def _import_reference():
    sys.path.insert(0, REPO)
    from tests.golden import ref_shims
    ref_shims.install(REF)
    os.chdir(REF)                      # BaseAviary.py:99 opens "Sol/resources/safegym/cf2x.urdf" relative to the cwd
    with contextlib.redirect_stdout(io.StringIO()):
        from Sol.Model.Environments.PBDroneEnv import PBDroneEnv
        from Sol.Model.Environments import normalize
        from Sol.PyBullet.enums import ActionType, Physics
        from Sol.Utilities import Waypoints

    class _DynEnv(PBDroneEnv):
        PHYSICS = property(lambda self: Physics.DYN, lambda self, value: None)        # see (2) above
        TIMESTEP = property(lambda self: self.PYB_TIMESTEP)                           # see (3) above

    return _DynEnv, normalize, ActionType, Physics, Waypoints


def make_env(ref, track, S, max_steps, normalize_obs):
    DynEnv, normalize, ActionType, Physics, Waypoints = ref
    if track == "circle":      # simulation_controller.py / PBDroneSimulator.py:111-130
        tr = Waypoints.Track(Waypoints.circle(radius=1, num_points=6, height=1), circle=True)
    else:
        tr = Waypoints.Track(Waypoints.reaching(), circle=False)
    targets = list(tr.waypoints)               # dilate_targets(.., 0) is the identity (PBDroneSimulator.py:89-105)
    if tr.is_circle:
        targets.pop(0)                         # PBDroneSimulator.py:129-130
    env = DynEnv(target_points=targets, threshold=0.3, discount=0.999, max_steps=max_steps, act=ActionType.THRUST,
                 gui=False, initial_xyzs=tr.initial_xyzs, save_folder=None, aviary_dim=tr.aviary_dim,
                 random_spawn=False, cylinder=True, circle=tr.is_circle, include_distance=True,
                 normalize_actions=True, collect_rollouts=False, physics=Physics.DYN,
                 pyb_freq=240, ctrl_freq=240 // S)       # PBDroneSimulator.py:154-172
    env.reset(seed=0)                                    # :173
    raw = env
    if normalize_obs:
        env = normalize.NormalizeObservation(env)        # :181
    return env, raw


def run(ref, track, S, mode, N, T, seed, max_steps, normalize_obs):
    with contextlib.redirect_stdout(io.StringIO()):
        envs = [make_env(ref, track, S, max_steps, normalize_obs) for _ in range(N)]
        a = actions(mode, T, N, seed)
        obs0 = np.stack([np.asarray(e.reset()[0], np.float64) for e, _ in envs])   # VecEnv.reset()
        D = obs0.shape[1]
        out = dict(actions=a, obs0=obs0, obs=np.zeros((T, N, D)), terminal_obs=np.full((T, N, D), np.nan),
                   reward=np.zeros((T, N)), done=np.zeros((T, N), np.uint8), found_targets=np.zeros((T, N), np.int32),
                   ep_return=np.full((T, N), np.nan), ep_length=np.zeros((T, N), np.int32),
                   pos=np.zeros((T, N, 3)), quat=np.zeros((T, N, 4)), vel=np.zeros((T, N, 3)),
                   rpy_rates=np.zeros((T, N, 3)), ang_v=np.zeros((T, N, 3)), dist=np.zeros((T, N)),
                   rpm=np.zeros((T, N, 4)))
        ep_ret, ep_len = np.zeros(N), np.zeros(N, np.int64)
        for t in range(T):
            for i, (e, raw) in enumerate(envs):
                o, r, term, trunc, info = e.step(a[t, i])
                ep_ret[i] += float(r)
                ep_len[i] += 1
                out["reward"][t, i] = float(r)
                out["done"][t, i] = (1 if term else 0) | (2 if trunc else 0)
                out["found_targets"][t, i] = info["found_targets"]
                # physical state right after the step (before any reset)
                out["pos"][t, i], out["quat"][t, i], out["vel"][t, i] = raw.pos[0], raw.quat[0], raw.vel[0]
                out["rpy_rates"][t, i], out["ang_v"][t, i] = raw.rpy_rates[0], raw.ang_v[0]
                out["dist"][t, i] = raw._distance_to_target
                out["rpm"][t, i] = raw.last_clipped_action[0]
                if term or trunc:
                    out["terminal_obs"][t, i] = o
                    out["ep_return"][t, i], out["ep_length"][t, i] = ep_ret[i], ep_len[i]
                    ep_ret[i], ep_len[i] = 0.0, 0
                    o, _ = e.reset()
                out["obs"][t, i] = o
    out["meta"] = np.array([track, str(S), mode, str(max_steps), "1" if normalize_obs else "0"])
    return out


if __name__ == "__main__":
    ref = _import_reference()
    for name, cfg in CASES.items():
        out = run(ref, *cfg)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "dones", int((out["done"] != 0).sum()), "truncs", int((out["done"] & 2).astype(bool).sum()),
              "max found", int(out["found_targets"].max()), file=sys.stderr)
