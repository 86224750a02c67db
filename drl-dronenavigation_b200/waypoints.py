"""Waypoint tracks: the data API of Sol/Utilities/Waypoints.py (Track :9-20 and the
generators :46-197) with the same signatures and return convention
``(waypoints, initial_xyzs, aviary_dim)``, minus the matplotlib imports, plus
``dilate_targets`` / ``track_targets`` (Sol/Model/PBDroneSimulator.py:89-105,129-130).
"""
from __future__ import annotations

import math

import numpy as np

_DIM_2M = (-2, -2, 0, 2, 2, 2)
_DIM_4M = (-4, -4, 0, 4, 4, 4)


class Track:
    """A track generator's result with the circle flag (Waypoints.py:9-20)."""

    def __init__(self, track, circle=False):
        self.waypoints, self.initial_xyzs, self.aviary_dim = track
        self.is_circle = circle

    def __str__(self):
        return (f"Track with {len(self.waypoints)} waypoints, initial position of: {self.initial_xyzs}, "
                f"and aviary dimensions of: {self.aviary_dim}.")


def normalize_coordinates(coordinates, new_size):
    """Affinely maps each axis of an (n, 3) array onto [0, new_size] (Waypoints.py:22-43)."""
    lo = coordinates.min(axis=0)
    span = coordinates.max(axis=0) - lo
    return (coordinates - lo) * (new_size / span)


def parametric_eq(num_points=5):
    """Closed sine/cosine curve sampled at num_points (Waypoints.py:46-58; y uses cos as there)."""
    theta = np.linspace(0, 2 * np.pi, num_points)
    radius = 5.0
    xs, ys, zs = radius * np.cos(theta), radius * np.cos(theta), 0.1 * np.sin(theta)
    return [np.array([xs[i], ys[i], zs[i]]) for i in range(num_points)]


def _pts(rows):
    return [np.array(r, dtype=np.float64) for r in rows]


def up():
    """Vertical climb (Waypoints.py:61-68)."""
    return (_pts([(0, 0, .1), (0, 0, .2), (0, 0, .5), (0, 0, .7), (0, 0, 1)]),
            np.array([0.0, 0.0, 0.1]), np.array(_DIM_2M))


def half_up_forward():
    """Climb then forward (Waypoints.py:71-78)."""
    return (_pts([(0, 0, .5), (0, 0, 1), (0, 1, 1.5)]), np.array([0., 0., 0.1]), np.array(_DIM_2M))


def up_circle():
    """Climbing loop back to the start (Waypoints.py:81-95)."""
    rows = [(0, 0, .2), (.1, 0, .3), (.1, .2, .7), (.3, .5, 1.5), (.5, 1, 1.5), (1, 1, 1.5),
            (1.5, 1, 1.5), (1.5, 1.5, 1), (1.5, .5, 1), (1, .5, .5), (.5, .2, .2), (0, 0, .2)]
    return _pts(rows), np.array([[0.0, 0.0, 0.1]]), np.array(_DIM_2M)


def up_sharp_back_turn():
    """Climb with a reversal (Waypoints.py:98-105)."""
    rows = [(0, 0, .5), (-.5, .2, .7), (.3, .5, .7), (1, .5, 1), (1.5, 1, 1.2)]
    return _pts(rows), np.array([[0.0, 0.0, 0.1]]), np.array(_DIM_2M)


def circle(radius, num_points, height, center=(0, 0, 0), plane="XY"):
    """num_points + 1 points on a circle (first == last) in the given plane, the spawn point
    (radius, 0, center_z + radius) and the 2 m box (Waypoints.py:108-139)."""
    ang = np.linspace(0, 2 * np.pi, num_points + 1, endpoint=True)
    pts = np.zeros((num_points + 1, 3))
    c, s = radius * np.cos(ang), radius * np.sin(ang)
    if plane == "XY":
        pts[:, 0], pts[:, 1], pts[:, 2] = center[0] + c, center[1] + s, center[2] + height
    elif plane == "XZ":
        pts[:, 0], pts[:, 2], pts[:, 1] = center[0] + c, center[2] + s + height, center[1]
    elif plane == "YZ":
        pts[:, 1], pts[:, 2], pts[:, 0] = center[1] + c, center[2] + s + height, center[0]
    else:
        raise ValueError("Invalid plane specified.")
    return pts, np.array([[radius, 0, center[2] + radius]]), np.array(_DIM_2M)


def generate_random_targets(num_targets: int) -> np.ndarray:
    """Random targets on a shell around the origin, z floored at 0.1 (Waypoints.py:142-169)."""
    targets = np.zeros((num_targets, 3))
    thetas = np.random.uniform(0.0, 2.0 * math.pi, size=(num_targets,))
    phis = np.random.uniform(0.0, 2.0 * math.pi, size=(num_targets,))
    for i in range(num_targets):
        dist = np.random.uniform(low=1.0, high=0.9)
        z = abs(dist * math.cos(phis[i]))
        targets[i] = (dist * math.sin(phis[i]) * math.cos(thetas[i]),
                      dist * math.sin(phis[i]) * math.sin(thetas[i]),
                      z if z > 0.1 else 0.1)
    return targets


def reaching():
    """The 7-gate race track of arXiv 2310.10943, closed, lifted by 3 and scaled by 1/5
    (Waypoints.py:172-197).  Spawn = gate 0."""
    gates = np.array([[-2.5, 4.5, 3], [10, 3.5, 1], [8, -4.5, 1], [-4.5, -6, 2],
                      [-5, -5, 2], [5, -1, 3], [2.5, 6, 3], [-2.5, 4.5, 3]], dtype=np.float64)
    gates[:, 2] += 3
    gates = gates / 5
    # the reference divides row by row after the z shift; numerically identical
    return gates, np.array([gates[0]]), np.array(_DIM_4M)


def dilate_targets(targets, factor: int) -> list:
    """Inserts `factor` evenly spaced points between consecutive targets
    (PBDroneSimulator.py:89-105)."""
    out = []
    for a, b in zip(targets[:-1], targets[1:]):
        out.extend(np.linspace(a, b, num=factor + 2)[:-1])
    out.append(targets[-1])
    return out


def track_targets(track: Track, target_factor: int = 0) -> list:
    """The target list PBDroneSimulator.__init__ hands to the env (:126-130): dilated, and for
    circle tracks without the first point (which is the spawn)."""
    targets = dilate_targets(track.waypoints, target_factor)
    if track.is_circle:
        targets.pop(0)
    return targets
