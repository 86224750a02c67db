"""CF2X constants and the URDF parameter reader.

Values: Sol/resources/safegym/cf2x.urdf:5,11-12,34 ; derived quantities follow
Sol/PyBullet/BaseAviary.py:76,163-176 ; the parser mirrors the attribute layout that
BaseAviary._parse_urdf_parameters (BaseAviary.py:1123-1163) reads.
The CUDA library carries its own copy of these numbers (csrc/dronenav.cu, struct CF2X);
tests/test_constants.py checks the two against each other through the action map.
"""
from __future__ import annotations

import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass

import numpy as np

RESOURCES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "resources")


@dataclass(frozen=True)
class DroneParameters:
    M: float
    L: float
    THRUST2WEIGHT_RATIO: float
    IXX: float
    IYY: float
    IZZ: float
    KF: float
    KM: float
    COLLISION_H: float
    COLLISION_R: float
    COLLISION_Z_OFFSET: float
    MAX_SPEED_KMH: float
    GND_EFF_COEFF: float
    PROP_RADIUS: float
    DRAG_COEFF_XY: float
    DRAG_COEFF_Z: float
    DW_COEFF_1: float
    DW_COEFF_2: float
    DW_COEFF_3: float
    PWM2RPM_SCALE: float
    PWM2RPM_CONST: float
    MIN_PWM: float
    MAX_PWM: float
    G: float = 9.8

    # ---- derived (BaseAviary.py:163-176) ----
    @property
    def J(self):
        return np.diag([self.IXX, self.IYY, self.IZZ])

    @property
    def J_INV(self):
        return np.linalg.inv(self.J)

    @property
    def DRAG_COEFF(self):
        return np.array([self.DRAG_COEFF_XY, self.DRAG_COEFF_XY, self.DRAG_COEFF_Z])

    @property
    def GRAVITY(self):
        return self.G * self.M

    @property
    def HOVER_RPM(self):
        return float(np.sqrt(self.GRAVITY / (4 * self.KF)))

    @property
    def MAX_RPM(self):
        return float(np.sqrt((self.THRUST2WEIGHT_RATIO * self.GRAVITY) / (4 * self.KF)))

    @property
    def MAX_THRUST(self):
        return 4 * self.KF * self.MAX_RPM ** 2

    @property
    def MAX_XY_TORQUE(self):
        return (2 * self.L * self.KF * self.MAX_RPM ** 2) / np.sqrt(2)

    @property
    def MAX_Z_TORQUE(self):
        return 2 * self.KM * self.MAX_RPM ** 2

    @property
    def GND_EFF_H_CLIP(self):
        return float(0.25 * self.PROP_RADIUS * np.sqrt(
            (15 * self.MAX_RPM ** 2 * self.KF * self.GND_EFF_COEFF) / self.MAX_THRUST))

    def physical_action_bounds(self):
        """Per-motor thrust bounds as float32 4-vectors (PBDroneEnv.py:113-116)."""
        lo = self.KF * (self.PWM2RPM_SCALE * self.MIN_PWM + self.PWM2RPM_CONST) ** 2
        hi = self.KF * (self.PWM2RPM_SCALE * self.MAX_PWM + self.PWM2RPM_CONST) ** 2
        return np.full(4, lo, np.float32), np.full(4, hi, np.float32)


def parse_urdf_parameters(file_name: str) -> DroneParameters:
    """Reads a drone URDF laid out like the reference's (properties element first, then
    the base link with inertial / collision children)."""
    root = ET.parse(file_name).getroot()
    props = root.find("properties").attrib
    base = root.find("link")
    inertial = base.find("inertial")
    inertia = inertial.find("inertia").attrib
    collision = base.find("collision")
    cyl = collision.find("geometry").find("cylinder").attrib
    offs = [float(v) for v in collision.find("origin").attrib["xyz"].split()]
    f = lambda k: float(props[k])
    return DroneParameters(
        M=float(inertial.find("mass").attrib["value"]), L=f("arm"), THRUST2WEIGHT_RATIO=f("thrust2weight"),
        IXX=float(inertia["ixx"]), IYY=float(inertia["iyy"]), IZZ=float(inertia["izz"]),
        KF=f("kf"), KM=f("km"),
        COLLISION_H=float(cyl["length"]), COLLISION_R=float(cyl["radius"]), COLLISION_Z_OFFSET=offs[2],
        MAX_SPEED_KMH=f("max_speed_kmh"), GND_EFF_COEFF=f("gnd_eff_coeff"), PROP_RADIUS=f("prop_radius"),
        DRAG_COEFF_XY=f("drag_coeff_xy"), DRAG_COEFF_Z=f("drag_coeff_z"),
        DW_COEFF_1=f("dw_coeff_1"), DW_COEFF_2=f("dw_coeff_2"), DW_COEFF_3=f("dw_coeff_3"),
        PWM2RPM_SCALE=f("pwm2rpm_scale"), PWM2RPM_CONST=f("pwm2rpm_const"),
        MIN_PWM=f("pwm_min"), MAX_PWM=f("pwm_max"))


CF2X = parse_urdf_parameters(os.path.join(RESOURCES, "cf2x.urdf"))
