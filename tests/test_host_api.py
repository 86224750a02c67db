"""Host-side mirror of the reference interface (no GPU): tracks, enums, constants, spaces, flags."""
import numpy as np
import pytest

import drl_dronenavigation_b200 as dn
from drl_dronenavigation_b200 import Waypoints
from drl_dronenavigation_b200.argparser import parse_args
from drl_dronenavigation_b200.spaces import Box


def test_circle_track_matches_reference_values():
    # Waypoints.circle(radius=1, num_points=6, height=1) -> 7 points, init [[1,0,1]], dim [-2,-2,0,2,2,2] (Waypoints.py:108-139)
    pts, init, dim = Waypoints.circle(radius=1, num_points=6, height=1)
    assert pts.shape == (7, 3) and np.allclose(pts[0], [1, 0, 1]) and np.allclose(pts[-1], [1, 0, 1], atol=1e-15)
    np.testing.assert_allclose(pts[1], [np.cos(np.pi / 3), np.sin(np.pi / 3), 1])
    np.testing.assert_array_equal(init, [[1, 0, 1]])
    np.testing.assert_array_equal(dim, [-2, -2, 0, 2, 2, 2])
    tr = dn.Track(Waypoints.circle(1, 6, 1), circle=True)
    t = dn.track_targets(tr)
    assert len(t) == 6 and t[-1][1] == pytest.approx(-2.45e-16, abs=1e-17)     # first point popped (PBDroneSimulator.py:129-130)
    with pytest.raises(ValueError):
        Waypoints.circle(1, 6, 1, plane="XX")


def test_reaching_track_and_oracle_agree():
    from oracle.dyn_oracle import circle_track, reaching_track
    g, init, dim = Waypoints.reaching()
    og, oinit, odim = reaching_track()
    np.testing.assert_array_equal(g, og)
    np.testing.assert_array_equal(init, oinit)
    np.testing.assert_array_equal(dim, odim)
    assert g.shape == (8, 3) and np.allclose(g[0], [-0.5, 0.9, 1.2]) and np.allclose(g[0], g[-1])
    ct, cinit, cdim = circle_track()
    np.testing.assert_array_equal(np.array(dn.track_targets(dn.Track(Waypoints.circle(1, 6, 1), circle=True))), ct)


def test_other_tracks_and_dilation():
    for f, n in ((Waypoints.up, 5), (Waypoints.half_up_forward, 3), (Waypoints.up_circle, 12), (Waypoints.up_sharp_back_turn, 5)):
        pts, init, dim = f()
        assert len(pts) == n and len(dim) == 6
    assert len(Waypoints.parametric_eq(5)) == 5
    d = dn.dilate_targets([np.zeros(3), np.ones(3), 2 * np.ones(3)], 1)
    np.testing.assert_allclose(np.array(d), [[0] * 3, [.5] * 3, [1] * 3, [1.5] * 3, [2] * 3])
    c = np.array([[0., 0, 0], [2, 4, 8]])
    np.testing.assert_allclose(Waypoints.normalize_coordinates(c, 1.0), [[0, 0, 0], [1, 1, 1]])


def test_enums_keep_reference_values():
    assert dn.ActionType.THRUST.value == "thrust" and dn.ActionType.RPM.value == "rpm"
    assert dn.Physics.DYN.value == "dyn" and dn.Physics.PYB_GND_DRAG_DW.value == "pyb_gnd_drag_dw"
    assert dn.DroneModel.CF2X.value == "cf2x" and dn.ObservationType.KIN.value == "kin"


def test_constants_from_urdf_match_oracle():
    from oracle.dyn_oracle import CF2X as O
    c = dn.CF2X
    for k in ("M", "L", "KF", "KM", "THRUST2WEIGHT_RATIO", "IXX", "IYY", "IZZ", "GND_EFF_COEFF", "PROP_RADIUS",
              "PWM2RPM_SCALE", "PWM2RPM_CONST", "MIN_PWM", "MAX_PWM", "COLLISION_H", "COLLISION_R"):
        assert getattr(c, k) == getattr(O, k), k
    assert c.HOVER_RPM == O.HOVER_RPM and c.GND_EFF_H_CLIP == O.GND_EFF_H_CLIP
    lo, hi = c.physical_action_bounds()
    assert lo[0] == np.float32(0.028161688) and hi[0] == np.float32(0.14834145)


def test_spaces_contract():
    from drl_dronenavigation_b200.vec_env import action_space, observation_space
    a = action_space(True)
    assert a.shape == (4,) and a.dtype == np.float32 and a.low.min() == -1 and a.high.max() == 1
    p = action_space(False)
    assert p.low[0] == np.float32(0.028161688)
    o = observation_space(True)
    assert o.shape == (13,) and o.low[2] == 0 and o.low[12] == 0 and o.high[12] == 1
    assert observation_space(False).shape == (12,)
    assert isinstance(o, Box) and a.contains(np.zeros(4, np.float32))


def test_flags_keep_reference_defaults():
    a = parse_args([])
    assert (a.num_envs, a.max_env_steps, a.agent, a.run_type, a.seed) == (12, 4096, "PPO", "full", 1)
    assert float(a.total_timesteps) == 10e6
    a = parse_args(["--agent", "PPO", "--run_type", "test", "--num_envs", "4096", "--savemodel", "f"])
    assert a.num_envs == 4096 and a.savemodel is False


def test_product_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from drl_dronenavigation_b200.batched_env import BatchedDroneEnv
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        BatchedDroneEnv(4, [[0, 0, 1]])
