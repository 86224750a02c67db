"""CPU ORACLE (test infrastructure, NOT product code).

A numpy restatement of the reference's per-environment control step for the
``Physics.DYN`` path of eRGiBi/DRL-DroneNavigation.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import this file; the product package never does.

PARITY PIN: the reference holds no golden vector, known-answer test or fixture
for this path, so the pin is the reference's own code run
here: ``tests/golden/make_ref_golden.py`` imports ``/root/reference`` UNMODIFIED
(behind import shims for the absent third-party packages, ``tests/golden/
ref_shims.py``) and steps ``PBDroneEnv`` / ``BaseAviary._dynamics`` /
``normalize.NormalizeObservation`` through nine scenarios; the committed outputs
(``tests/golden/ref_*.npz``) are reproduced by this file to FP64 round-off
(``tests/test_ref_golden.py``: observations identical as float32, reward 1e-12,
state 1e-13, done bits / found_targets / episode lengths exact).  Two things the
generator has to supply because the reference's DYN branch is unreachable as
shipped: ``Sol/PyBullet/BaseAviary.py:418`` forces ``Physics.PYB`` (neutralised by a
read-only ``PHYSICS`` property in a subclass) and ``BaseAviary.py:944`` reads the
undefined ``self.TIMESTEP`` (defined as ``PYB_TIMESTEP``).  What remains unpinned
is third-party: ``pybullet`` itself (unpinned by the reference, not installable
here) -- its three quaternion helpers are restated from the published bullet3
algorithms both in the shim and below.  The analytic known-answer tests in
``tests/test_oracle_kat.py`` and the oracle-minted fixtures
(``tests/golden/make_golden.py``: drag / ground-effect extension) complement it.

What each piece follows (all paths relative to ``/root/reference``):

* constants                 ``Sol/resources/safegym/cf2x.urdf:5,11-12,34``;
                            ``Sol/PyBullet/BaseAviary.py:76,84-85,163-176``
* action map (THRUST)       ``Sol/Model/Environments/PBDroneEnv.py:872-895,949-971``;
                            ``Sol/Model/env_utils.py:8-59``
* action map (RPM)          ``Sol/PyBullet/BaseSingleAgentAviary.py:176-179,211-212``
* action map (PID family)   ``Sol/PyBullet/BaseSingleAgentAviary.py:72-75,91,180-223``; ``Sol/PyBullet/BaseAviary.py:1255-1297``;
                            ``Sol/PyBullet/DSLPIDControl.py:20-261``; ``Sol/PyBullet/BaseControl.py:20-52``.  scipy's
                            ``Rotation.from_matrix(..).as_euler('XYZ')`` / ``from_euler('XYZ', ..)`` (third party, present in
                            this image but kept out of the oracle) are restated in numpy; the fixtures
                            ``tests/golden/ref_*_{pid,vel,one_d_pid}*.npz`` were minted with the real scipy
* airframes                 ``Sol/resources/cf2p.urdf:5,11-12,34,42-78``; ``Sol/resources/racer.urdf:5,11-12,28,36-72``;
                            torque-mix branches ``Sol/PyBullet/BaseAviary.py:927-935``
* substep loop / ordering   ``Sol/PyBullet/BaseAviary.py:407-453``
* rigid body                ``Sol/PyBullet/BaseAviary.py:899-973``
* drag / ground effect      ``Sol/PyBullet/BaseAviary.py:838-865,798-834`` (formulas only;
                            "DYN + drag + ground effect" is our documented extension)
* observation               ``PBDroneEnv.py:296-398``
* reward                    ``PBDroneEnv.py:475-607``
* termination / truncation  ``PBDroneEnv.py:444-473,678-786``
* post-step / reset         ``PBDroneEnv.py:201-223,609-665``; ``BaseAviary.py:276-320,527-598``
* wrappers                  ``Sol/Model/Environments/normalize.py:10-147`` and the SB3
                            ``Monitor`` / ``SubprocVecEnv`` worker auto-reset contract
* third-party pybullet math restated from the published bullet3 algorithms
  (``btMatrix3x3::setRotation`` and ``pybullet.c:getEulerFromQuaternion``);
  pybullet is unpinned by the reference (absent from requirements.txt / uv.lock).

Arithmetic follows the reference dtype-for-dtype: the THRUST action path and
the per-motor force / z-torque products are float32 (numpy keeps float32 for
float32-array x python-scalar), everything downstream is float64, the
observation is cast to float32 at the end.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

# --------------------------------------------------------------------------
# constants
# --------------------------------------------------------------------------


@dataclass(frozen=True)
class CF2XConstants:
    """CF2X physical constants (safegym/cf2x.urdf:5,11-12,34) and the values
    BaseAviary derives from them (BaseAviary.py:76,163-176)."""

    M: float = 0.027
    L: float = 0.0397
    THRUST2WEIGHT_RATIO: float = 2.25
    IXX: float = 1.4e-5
    IYY: float = 1.4e-5
    IZZ: float = 2.17e-5
    KF: float = 3.16e-10
    KM: float = 7.94e-12
    COLLISION_H: float = 0.025
    COLLISION_R: float = 0.06
    COLLISION_Z_OFFSET: float = 0.0
    MAX_SPEED_KMH: float = 30.0
    GND_EFF_COEFF: float = 11.36859
    PROP_RADIUS: float = 2.31348e-2
    DRAG_COEFF_XY: float = 9.1785e-7
    DRAG_COEFF_Z: float = 10.311e-7
    DW_COEFF_1: float = 2267.18
    DW_COEFF_2: float = 0.16
    DW_COEFF_3: float = -0.11
    PWM2RPM_SCALE: float = 0.2685
    PWM2RPM_CONST: float = 4070.3
    MIN_PWM: float = 20000.0
    MAX_PWM: float = 65535.0
    G: float = 9.8
    # safegym prop layout (safegym/cf2x.urdf:42,54,66,78), used by the ground-effect extension only.
    # It is the layout the DYN torque mix (BaseAviary.py:931-932) is consistent with:
    # tau_x = sum(y_i f_i) = (f0+f1-f2-f3) d, tau_y = -sum(x_i f_i) = (-f0+f1+f2-f3) d.
    PROP_XY: tuple = ((0.028, 0.028), (-0.028, 0.028), (-0.028, -0.028), (0.028, -0.028))

    @property
    def GRAVITY(self) -> float:
        return self.G * self.M

    @property
    def HOVER_RPM(self) -> float:
        return float(np.sqrt(self.GRAVITY / (4 * self.KF)))

    @property
    def MAX_RPM(self) -> float:
        return float(np.sqrt((self.THRUST2WEIGHT_RATIO * self.GRAVITY) / (4 * self.KF)))

    @property
    def MAX_THRUST(self) -> float:
        return 4 * self.KF * self.MAX_RPM ** 2

    @property
    def GND_EFF_H_CLIP(self) -> float:
        return float(0.25 * self.PROP_RADIUS * np.sqrt(
            (15 * self.MAX_RPM ** 2 * self.KF * self.GND_EFF_COEFF) / self.MAX_THRUST))

    @property
    def J(self) -> np.ndarray:
        return np.diag([self.IXX, self.IYY, self.IZZ])

    @property
    def J_INV(self) -> np.ndarray:
        return np.linalg.inv(self.J)


CF2X = CF2XConstants()
# The other two airframes of Sol/PyBullet/enums.py:3-8.  Their URDFs carry no pwm attributes (the values below are the
# class defaults, only ever used by DSLPIDControl, which hard-codes the same numbers: DSLPIDControl.py:43-46).
CF2P = CF2XConstants(IXX=2.3951e-5, IYY=2.3951e-5, IZZ=3.2347e-5,
                     PROP_XY=((0.0397, 0.0), (0.0, 0.0397), (-0.0397, 0.0), (0.0, -0.0397)))
RACE = CF2XConstants(M=0.830, L=0.109, THRUST2WEIGHT_RATIO=4.17, IXX=3.113e-3, IYY=3.113e-3, IZZ=3.113e-3,
                     KF=8.47e-9, KM=2.13e-11, MAX_SPEED_KMH=200.0, PROP_RADIUS=12.7e-2,
                     PROP_XY=((0.085, 0.0675), (-0.085, 0.0675), (-0.085, -0.0675), (0.085, -0.0675)))
MODEL_CF2X, MODEL_CF2P, MODEL_RACE = "cf2x", "cf2p", "racer"
AIRFRAMES = {MODEL_CF2X: CF2X, MODEL_CF2P: CF2P, MODEL_RACE: RACE}


# --------------------------------------------------------------------------
# pybullet math helpers restated (bullet3: LinearMath/btMatrix3x3.h setRotation;
# examples/pybullet/pybullet.c pybullet_getEulerFromQuaternion)
# --------------------------------------------------------------------------


def bullet_matrix_from_quaternion(q) -> np.ndarray:
    """p.getMatrixFromQuaternion (BaseAviary.py:920): btMatrix3x3::setRotation."""
    x, y, z, w = (float(q[0]), float(q[1]), float(q[2]), float(q[3]))
    d = x * x + y * y + z * z + w * w
    s = 2.0 / d
    xs, ys, zs = x * s, y * s, z * s
    wx, wy, wz = w * xs, w * ys, w * zs
    xx, xy, xz = x * xs, x * ys, x * zs
    yy, yz, zz = y * ys, y * zs, z * zs
    return np.array([
        [1.0 - (yy + zz), xy - wz, xz + wy],
        [xy + wz, 1.0 - (xx + zz), yz - wx],
        [xz - wy, yz + wx, 1.0 - (xx + yy)],
    ])


def bullet_euler_from_quaternion(q) -> np.ndarray:
    """p.getEulerFromQuaternion (BaseAviary.py:597)."""
    x, y, z, w = (float(q[0]), float(q[1]), float(q[2]), float(q[3]))
    sqx, sqy, sqz, squ = x * x, y * y, z * z, w * w
    sarg = -2.0 * (x * z - w * y)
    if sarg <= -0.99999:
        return np.array([0.0, -0.5 * math.pi, 2.0 * math.atan2(x, -y)])
    if sarg >= 0.99999:
        return np.array([0.0, 0.5 * math.pi, 2.0 * math.atan2(-x, y)])
    return np.array([
        math.atan2(2.0 * (y * z + w * x), squ - sqx - sqy + sqz),
        math.asin(sarg),
        math.atan2(2.0 * (x * y + w * z), squ + sqx - sqy - sqz),
    ])


def bullet_quaternion_from_euler(rpy) -> np.ndarray:
    """p.getQuaternionFromEuler (BaseAviary.py:567): btQuaternion::setEulerZYX."""
    r, p_, y = (float(rpy[0]) * 0.5, float(rpy[1]) * 0.5, float(rpy[2]) * 0.5)
    cr, sr = math.cos(r), math.sin(r)
    cp, sp = math.cos(p_), math.sin(p_)
    cy, sy = math.cos(y), math.sin(y)
    return np.array([
        sr * cp * cy - cr * sp * sy,
        cr * sp * cy + sr * cp * sy,
        cr * cp * sy - sr * sp * cy,
        cr * cp * cy + sr * sp * sy,
    ])


def bullet_pose_readback(q) -> np.ndarray:
    """resetBasePositionAndOrientation -> getBasePositionAndOrientation
    (BaseAviary.py:946-950 then :596) goes through a btTransform, so the
    quaternion that comes back is unit length (its sign may flip, which no
    downstream quantity can observe: R(q), the Euler angles and _integrateQ are
    all even / linear in q).  Restated as a plain normalisation."""
    q = np.asarray(q, dtype=np.float64)
    return q / math.sqrt(float(np.dot(q, q)))


# --------------------------------------------------------------------------
# action maps
# --------------------------------------------------------------------------


def physical_action_bounds(c: CF2XConstants = CF2X):
    """PBDroneEnv.py:113-116 (float32 arrays)."""
    a_low = c.KF * (c.PWM2RPM_SCALE * c.MIN_PWM + c.PWM2RPM_CONST) ** 2
    a_high = c.KF * (c.PWM2RPM_SCALE * c.MAX_PWM + c.PWM2RPM_CONST) ** 2
    return np.full(4, a_low, np.float32), np.full(4, a_high, np.float32)


def rescale_action(action, bounds):
    """PBDroneEnv.py:949-971.  NB: maps physical -> normalised although it is fed
    a normalised action; reproduced as is."""
    min_action = np.zeros(4, dtype=np.float32) + bounds[0]
    max_action = np.zeros(4, dtype=np.float32) + bounds[1]
    low = -1 * np.ones(4, dtype=np.float32)
    high = np.ones(4, dtype=np.float32)
    out = low + (high - low) * ((action - min_action) / (max_action - min_action))
    return np.clip(out, low, high)


def thrust_to_rpm(action, bounds, c: CF2XConstants = CF2X):
    """PBDroneEnv.py:872-895 + env_utils.py:8-59 for a 4-vector (n_motor = 1)."""
    thrust = np.clip(action, bounds[0], bounds[1])
    n_motor = 4 // int(thrust.size)
    thrust = np.clip(thrust, np.zeros_like(thrust), None)
    pwm = (np.sqrt(thrust / n_motor / c.KF) - c.PWM2RPM_CONST) / c.PWM2RPM_SCALE
    pwm = np.clip(np.array(pwm), c.MIN_PWM, c.MAX_PWM)
    return c.PWM2RPM_SCALE * pwm + c.PWM2RPM_CONST


def thrust_to_rpm_batch(actions, bounds, c: CF2XConstants = CF2X):
    """thrust_to_rpm for many 4-vectors at once ([M, 4] float32): the same float32 operations with
    n_motor = 1 written out (the literal function derives n_motor from the array size, so it only
    accepts one 4-vector).  tests/test_oracle_kat.py checks the two agree bit for bit."""
    a = np.asarray(actions, dtype=np.float32).reshape(-1, 4)
    thrust = np.clip(a, bounds[0], bounds[1])
    thrust = np.clip(thrust, np.zeros_like(thrust), None)
    pwm = (np.sqrt(thrust / 1 / c.KF) - c.PWM2RPM_CONST) / c.PWM2RPM_SCALE
    pwm = np.clip(pwm, c.MIN_PWM, c.MAX_PWM)
    return c.PWM2RPM_SCALE * pwm + c.PWM2RPM_CONST


def rescale_action_batch(actions, bounds):
    """rescale_action for [M, 4] float32 (numpy broadcasting makes the literal function batch-safe)."""
    return rescale_action(np.asarray(actions, dtype=np.float32).reshape(-1, 4), bounds)


def rpm_action_to_rpm(action, c: CF2XConstants = CF2X, numpy_legacy_cast: bool = True):
    """BaseSingleAgentAviary.py:176-179.  Under the reference's pinned numpy 1.26
    (uv.lock:231-232) ``np.float64 scalar * float32 array`` stays float32
    (value-based casting); numpy >= 2 would promote to float64.  The pinned
    behaviour is the default."""
    a = np.asarray(action)
    if numpy_legacy_cast and a.dtype == np.float32:
        one_plus = np.float32(1) + np.float32(0.05) * a
        return np.float32(c.HOVER_RPM) * one_plus
    # numpy >= 2 (NEP 50): the float32 array (1 + 0.05 a) times the np.float64 SCALAR HOVER_RPM (BaseAviary.py:164) is float64
    return np.array(np.float64(c.HOVER_RPM) * (1 + 0.05 * a), dtype=np.float64)


# --------------------------------------------------------------------------
# DSLPIDControl (Sol/PyBullet/DSLPIDControl.py:20-261 over BaseControl.py:20-52)
# --------------------------------------------------------------------------


def euler_XYZ_from_matrix(R):
    """scipy ``Rotation.from_matrix(R).as_euler('XYZ')`` (intrinsic x-y'-z''), DSLPIDControl.py:193:
    R = Rx(a) Ry(b) Rz(c) -> b = asin(R02), a = atan2(-R12, R22), c = atan2(-R01, R00) (away from gimbal lock)."""
    return np.array([math.atan2(-R[1, 2], R[2, 2]), math.asin(min(1.0, max(-1.0, R[0, 2]))), math.atan2(-R[0, 1], R[0, 0])])


def matrix_from_euler_XYZ(e):
    """scipy ``Rotation.from_euler('XYZ', e)`` -> ``as_quat`` -> ``from_quat`` -> ``as_matrix`` (DSLPIDControl.py:233-235;
    the ``w,x,y,z = target_quat`` / ``from_quat([w, x, y, z])`` pair re-assembles the same array, a no-op)."""
    ca, sa, cb, sb, cc, sc = math.cos(e[0]), math.sin(e[0]), math.cos(e[1]), math.sin(e[1]), math.cos(e[2]), math.sin(e[2])
    return np.array([[cb * cc, -cb * sc, sb],
                     [ca * sc + sa * sb * cc, ca * cc - sa * sb * sc, -sa * cb],
                     [sa * sc - ca * sb * cc, sa * cc + ca * sb * sc, ca * cb]])


class OracleDSLPID:
    """DSLPIDControl(drone_model=DroneModel.CF2X): BaseSingleAgentAviary.py:72-73 constructs the CF2X controller for CF2X
    and CF2P airframes alike, and never resets it (no ``ctrl.reset()`` in any env reset)."""

    def __init__(self, g=9.8, c: CF2XConstants = CF2X):
        self.GRAVITY = g * c.M                      # BaseControl.py:33 (URDF of the controller's model: cf2x)
        self.KF = c.KF
        self.P_COEFF_FOR = np.array([.4, .4, 1.25])
        self.I_COEFF_FOR = np.array([.05, .05, .05])
        self.D_COEFF_FOR = np.array([.2, .2, .5])
        self.P_COEFF_TOR = np.array([70000., 70000., 60000.])
        self.I_COEFF_TOR = np.array([.0, .0, 500.])
        self.D_COEFF_TOR = np.array([20000., 20000., 12000.])
        self.PWM2RPM_SCALE, self.PWM2RPM_CONST, self.MIN_PWM, self.MAX_PWM = 0.2685, 4070.3, 20000, 65535
        self.MIXER_MATRIX = np.array([[-.5, -.5, -1], [-.5, .5, 1], [.5, .5, -1], [.5, -.5, 1]])
        self.reset()

    def reset(self):                                # DSLPIDControl.py:66-80
        self.control_counter = 0
        self.last_rpy = np.zeros(3)
        self.integral_pos_e = np.zeros(3)
        self.integral_rpy_e = np.zeros(3)

    def computeControl(self, control_timestep, cur_pos, cur_quat, cur_vel, cur_ang_vel, target_pos,
                       target_rpy=np.zeros(3), target_vel=np.zeros(3), target_rpy_rates=np.zeros(3)):
        self.control_counter += 1
        thrust, computed_target_rpy, pos_e = self._dslPIDPositionControl(
            control_timestep, cur_pos, cur_quat, cur_vel, target_pos, target_rpy, target_vel)
        rpm = self._dslPIDAttitudeControl(control_timestep, thrust, cur_quat, computed_target_rpy, target_rpy_rates)
        cur_rpy = bullet_euler_from_quaternion(cur_quat)
        return rpm, pos_e, computed_target_rpy[2] - cur_rpy[2]

    def _dslPIDPositionControl(self, control_timestep, cur_pos, cur_quat, cur_vel, target_pos, target_rpy, target_vel):
        cur_rotation = bullet_matrix_from_quaternion(cur_quat)                                # :163
        pos_e = target_pos - cur_pos
        vel_e = target_vel - cur_vel
        self.integral_pos_e = self.integral_pos_e + pos_e * control_timestep
        self.integral_pos_e = np.clip(self.integral_pos_e, -2., 2.)
        self.integral_pos_e[2] = np.clip(self.integral_pos_e[2], -0.15, .15)
        target_thrust = (np.multiply(self.P_COEFF_FOR, pos_e) + np.multiply(self.I_COEFF_FOR, self.integral_pos_e)
                         + np.multiply(self.D_COEFF_FOR, vel_e) + np.array([0, 0, self.GRAVITY]))
        scalar_thrust = max(0., np.dot(target_thrust, cur_rotation[:, 2]))
        thrust = (math.sqrt(scalar_thrust / (4 * self.KF)) - self.PWM2RPM_CONST) / self.PWM2RPM_SCALE
        target_z_ax = target_thrust / np.linalg.norm(target_thrust)
        target_x_c = np.array([math.cos(target_rpy[2]), math.sin(target_rpy[2]), 0])
        target_y_ax = np.cross(target_z_ax, target_x_c) / np.linalg.norm(np.cross(target_z_ax, target_x_c))
        target_x_ax = np.cross(target_y_ax, target_z_ax)
        target_rotation = (np.vstack([target_x_ax, target_y_ax, target_z_ax])).transpose()
        target_euler = euler_XYZ_from_matrix(target_rotation)                                 # :193
        return thrust, target_euler, pos_e

    def _dslPIDAttitudeControl(self, control_timestep, thrust, cur_quat, target_euler, target_rpy_rates):
        cur_rotation = bullet_matrix_from_quaternion(cur_quat)
        cur_rpy = np.array(bullet_euler_from_quaternion(cur_quat))
        target_rotation = matrix_from_euler_XYZ(target_euler)                                 # :233-235
        rot_matrix_e = np.dot(target_rotation.transpose(), cur_rotation) - np.dot(cur_rotation.transpose(), target_rotation)
        rot_e = np.array([rot_matrix_e[2, 1], rot_matrix_e[0, 2], rot_matrix_e[1, 0]])
        rpy_rates_e = target_rpy_rates - (cur_rpy - self.last_rpy) / control_timestep
        self.last_rpy = cur_rpy
        self.integral_rpy_e = self.integral_rpy_e - rot_e * control_timestep
        self.integral_rpy_e = np.clip(self.integral_rpy_e, -1500., 1500.)
        self.integral_rpy_e[0:2] = np.clip(self.integral_rpy_e[0:2], -1., 1.)
        target_torques = (- np.multiply(self.P_COEFF_TOR, rot_e) + np.multiply(self.D_COEFF_TOR, rpy_rates_e)
                          + np.multiply(self.I_COEFF_TOR, self.integral_rpy_e))
        target_torques = np.clip(target_torques, -3200, 3200)
        pwm = thrust + np.dot(self.MIXER_MATRIX, target_torques)
        pwm = np.clip(pwm, self.MIN_PWM, self.MAX_PWM)
        return self.PWM2RPM_SCALE * pwm + self.PWM2RPM_CONST


def calculate_next_step(current_position, destination, step_size=1):
    """BaseAviary._calculateNextStep, BaseAviary.py:1255-1297."""
    direction = destination - current_position
    distance = np.linalg.norm(direction)
    if distance <= step_size:
        return destination
    return current_position + direction / distance * step_size


# --------------------------------------------------------------------------
# the environment (PBDroneEnv + BaseAviary, Physics.DYN)
# --------------------------------------------------------------------------

PHYSICS_DYN = "dyn"
PHYSICS_DYN_DRAG = "dyn_drag"
PHYSICS_DYN_GND = "dyn_gnd"
PHYSICS_DYN_GND_DRAG = "dyn_gnd_drag"

# --------------------------------------------------------------------------
# random spawn (optional; PBDroneEnv.py:622-629 is commented out in the reference)
# --------------------------------------------------------------------------


def philox4x32_10(counter, key):
    """Philox4x32-10 (Salmon et al., SC'11).  counter: 4 uint32, key: 2 uint32 -> 4 uint32."""
    M0, M1, W0, W1, MASK = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF
    c = [int(v) & MASK for v in counter]
    k0, k1 = int(key[0]) & MASK, int(key[1]) & MASK
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c[3] ^ k1) & MASK, p0 & MASK]
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c


def spawn_line(targets, aviary_dim, seed, global_env_id, counter, max_distance=0.1):
    """PositionGenerator.generate_random_point_around_line (position_generator.py:121-152) between two distinct
    random targets (PBDroneEnv.py:623-624), with the draws taken from Philox(seed; counter, block, env id)
    instead of numpy's / random's global generators (the mapping of draws to variables is ours; it is the same
    in csrc/dn_device.cuh::spawn_line)."""
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    gid = (global_env_id & 0xFFFFFFFF, (global_env_id >> 32) & 0xFFFFFFFF)
    a = philox4x32_10((counter, 0, gid[0], gid[1]), key)
    b = philox4x32_10((counter, 1, gid[0], gid[1]), key)
    u01 = lambda x: float(np.float32(x >> 8) * np.float32(1.0 / 16777216.0))
    u01_open = lambda x: float((np.float32(x >> 8) + np.float32(1.0)) * np.float32(1.0 / 16777216.0))
    T = len(targets)
    ia = min(int(np.float32(u01(a[0])) * np.float32(T)), T - 1)
    ib = min(int(np.float32(u01(a[1])) * np.float32(T - 1)), T - 2)
    if ib >= ia:
        ib += 1
    from_point, to_point = np.asarray(targets[ia], np.float64), np.asarray(targets[ib], np.float64)
    t = u01(a[2])
    point = from_point + t * (to_point - from_point)
    direction_vector = to_point - from_point
    r1, r2 = math.sqrt(-2.0 * math.log(u01_open(a[3]))), math.sqrt(-2.0 * math.log(u01_open(b[1])))
    random_vector = np.array([r1 * math.cos(2 * math.pi * u01(b[0])), r1 * math.sin(2 * math.pi * u01(b[0])),
                              r2 * math.cos(2 * math.pi * u01(b[2]))])
    perpendicular_vector = np.cross(direction_vector, random_vector)
    n = np.linalg.norm(perpendicular_vector)
    # identical end points (the reaching track's first and last gate coincide) make this 0/0 in the reference;
    # defined here, and in the kernel, as "no perpendicular offset"
    perpendicular_vector = perpendicular_vector / n if n > 0 else np.zeros(3)
    offset = (2.0 * u01(b[3]) - 1.0) * max_distance
    point = point + offset * perpendicular_vector
    lo, hi = np.asarray(aviary_dim[:3], np.float64), np.asarray(aviary_dim[3:], np.float64)
    return np.minimum(np.maximum(point, lo), hi)


def spawn_midpoint(targets, seed, global_env_id, counter):
    """PBDroneEnv.py:641-648 (commented out in the reference): a random segment of the track, its midpoint, and the
    target order rolled to start behind it.  Returns (spawn point, roll)."""
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    a = philox4x32_10((counter, 0, global_env_id & 0xFFFFFFFF, (global_env_id >> 32) & 0xFFFFFFFF), key)
    T = len(targets)
    u = np.float32(a[0] >> 8) * np.float32(1.0 / 16777216.0)
    segment_index = min(int(u * np.float32(T - 1)), T - 2)                     # np.random.randint(T - 1)
    segment_center = (np.asarray(targets[segment_index], np.float64) + np.asarray(targets[segment_index + 1], np.float64)) / 2
    return segment_center, segment_index + 1


ACT_THRUST = "thrust"
ACT_RPM = "rpm"
ACT_ONE_D_RPM = "one_d_rpm"
ACT_PID = "pid"
ACT_VEL = "vel"
ACT_ONE_D_PID = "one_d_pid"


class OracleDroneEnv:
    """Single-environment restatement.  Attribute names follow the reference so
    the code can be read side by side with PBDroneEnv.py / BaseAviary.py."""

    def __init__(self, target_points, threshold, discount, max_steps, aviary_dim,
                 initial_xyzs=None, initial_rpys=None, physics=PHYSICS_DYN,
                 pyb_freq=240, ctrl_freq=240, act=ACT_THRUST, cylinder=True,
                 circle=False, include_distance=False, normalize_actions=False,
                 ground_contact=False, reward_id="default", random_spawn=False, seed=0, global_env_id=0,
                 consts: CF2XConstants = None, drone_model=MODEL_CF2X):
        self.DRONE_MODEL = drone_model
        # RPM action types: float32 arithmetic as under the reference's pinned numpy 1.26 (see rpm_action_to_rpm); set to
        # False to reproduce fixtures minted under numpy >= 2 (float64 promotion)
        self.numpy_legacy_cast = True
        consts = AIRFRAMES[drone_model] if consts is None else consts
        self.C = consts
        if act in (ACT_PID, ACT_VEL, ACT_ONE_D_PID):           # BaseSingleAgentAviary.py:70-75,90-91
            if drone_model not in (MODEL_CF2X, MODEL_CF2P):
                raise ValueError("no controller is available for the specified drone_model")
            self.ctrl = OracleDSLPID()
            self.SPEED_LIMIT = 0.03 * consts.MAX_SPEED_KMH * (1000 / 3600)
        self.random_spawn, self.seed, self.global_env_id, self._reset_counter = random_spawn, seed, global_env_id, 0
        self.reward_id = reward_id        # which reference reward function runs inside the PBDroneEnv step machine
        self.EPISODE_LEN_SEC = 1          # PBDroneEnv.py:68 says 5, but _clipAndNormalizeState overwrites it with 1 on every
                                          # observation (PBDroneEnv.py:348), i.e. before the first reward is ever computed
        self.ACT_TYPE = act
        self.PHYSICS = physics
        self._target_points = np.array(target_points, dtype=np.float64)
        self._or_target_points = self._target_points.copy()          # PBDroneEnv.py:73
        self._threshold = threshold
        self._discount = discount
        self._max_steps = max_steps
        self._aviary_dim = aviary_dim
        (self._x_low, self._y_low, self._z_low,
         self._x_high, self._y_high, self._z_high) = [float(v) for v in aviary_dim]
        self.circle_radius = 1            # PBDroneEnv.py:84 (hard-coded)
        self.cylinder = cylinder
        self.circle = circle
        self.include_distance = include_distance
        self._max_target_dist = max(abs(self._x_low) + self._x_high,
                                    abs(self._y_low) + self._y_high, self._z_high)
        self.ground_contact = ground_contact  # documented extension; off = DYN (no contacts)

        # BaseAviary.__init__ : frequencies (BaseAviary.py:79-85)
        if pyb_freq % ctrl_freq != 0:
            raise ValueError("pyb_freq is not divisible by ctrl_freq")
        self.PYB_FREQ, self.CTRL_FREQ = pyb_freq, ctrl_freq
        self.PYB_STEPS_PER_CTRL = int(pyb_freq / ctrl_freq)
        self.PYB_TIMESTEP = 1.0 / pyb_freq
        self.CTRL_TIMESTEP = 1.0 / ctrl_freq
        c = consts
        self.GRAVITY = c.GRAVITY
        self.HOVER_RPM = c.HOVER_RPM
        self.J, self.J_INV = c.J, c.J_INV
        self.DRAG_COEFF = np.array([c.DRAG_COEFF_XY, c.DRAG_COEFF_XY, c.DRAG_COEFF_Z])

        if initial_xyzs is None:  # BaseAviary.py:248-253
            initial_xyzs = np.array([[0.0, 0.0, c.COLLISION_H / 2 - c.COLLISION_Z_OFFSET + .1]])
        self.INIT_XYZS = np.array(initial_xyzs, dtype=np.float64).reshape(1, 3)
        self.INIT_RPYS = (np.zeros((1, 3)) if initial_rpys is None
                          else np.array(initial_rpys, dtype=np.float64).reshape(1, 3))

        self.physical_action_bounds = physical_action_bounds(c)
        self.normalize_actions = normalize_actions

        self._housekeeping()

        # PBDroneEnv.__init__ tail (PBDroneEnv.py:122-144)
        self._current_position = self.INIT_XYZS[0].copy()
        self.current_vel, self.current_ang_v = np.zeros(3), np.zeros(3)
        self.prev_vel, self.prev_ang_v = np.zeros(3), np.zeros(3)
        self._distance_to_target = np.linalg.norm(self._current_position - self._target_points[0])
        self._prev_distance_to_target = np.linalg.norm(self._current_position - self._target_points[0])
        self._current_target_index = 0
        self.just_found = False
        self._is_done = False
        self._steps = 0
        self._last_position = self._current_position.copy()   # dummy_env.py:125 (reaching-progress reward only)
        self._last_action = np.zeros(4, dtype=np.float32)      # PBDroneEnv.py:131; never reset (:655 is commented out)
        # test instrumentation: signed distances of this step's threshold comparisons to their
        # thresholds (so FP32-vs-FP64 near-ties can be recognised); never read by the step itself
        self.margins = []

    # ---- BaseAviary._housekeeping (BaseAviary.py:527-584), DYN part --------
    def _housekeeping(self):
        self.step_counter = 0
        self.last_clipped_action = np.zeros(4)
        self.pos = self.INIT_XYZS[0].astype(np.float64).copy()
        self.quat = bullet_pose_readback(bullet_quaternion_from_euler(self.INIT_RPYS[0]))
        self.rpy = bullet_euler_from_quaternion(self.quat)
        self.vel = np.zeros(3)
        self.ang_v = np.zeros(3)
        self.rpy_rates = np.zeros(3)

    # ---- spaces -------------------------------------------------------------
    @property
    def obs_dim(self):
        return 13 if self.include_distance else 12

    # ---- reset (BaseAviary.py:276-320 then PBDroneEnv.py:609-665) ----------
    def reset(self, seed=None, options=None):
        self._reset_counter += 1
        if self.random_spawn:
            # OUR semantics for the reference's commented-out blocks (PBDroneEnv.py:622-629 / :641-648): the point is
            # drawn first, the episode starts there, the reset observation shows it and the distances are measured from it
            if self.random_spawn == "midpoint":
                center, roll = spawn_midpoint(self._or_target_points, self.seed, self.global_env_id, self._reset_counter)
                self.INIT_XYZS[0] = center
                self._target_points = np.concatenate((self._or_target_points[roll:], self._or_target_points[:roll]), axis=0)
            else:
                self.INIT_XYZS[0] = spawn_line(self._target_points, self._aviary_dim, self.seed, self.global_env_id, self._reset_counter)
            self._current_position = self.INIT_XYZS[0].copy()
        self._housekeeping()
        initial_obs = self._computeObs()      # BEFORE the distances are reset
        initial_info = self._computeInfo()    # found_targets of the previous episode
        self._is_done = False
        self._current_target_index = 0
        self._steps = 0
        self._distance_to_target = np.linalg.norm(self._current_position - self._target_points[0])
        self._prev_distance_to_target = np.linalg.norm(self._current_position - self._target_points[0])
        self.prev_vel, self.prev_ang_v = np.zeros(3), np.zeros(3)
        self.current_vel, self.current_ang_v = np.zeros(3), np.zeros(3)
        self.just_found = False
        return initial_obs, initial_info

    # ---- step (PBDroneEnv.py:171-199 around BaseAviary.py:324-453) ---------
    def step(self, action):
        action = np.asarray(action)
        self.margins = []
        self._cur_action = action            # the raw policy action of this step (literature rewards: a_t)
        self._pos_at_entry = self.pos.copy()
        a = self.rescale_action(action) if self.normalize_actions else action
        rpm = np.reshape(self._preprocessAction(a), 4)
        for _ in range(self.PYB_STEPS_PER_CTRL):
            # kinematics are (re)read from Bullet before every substep when S > 1
            # (BaseAviary.py:413-415) and after the last one (:444); with the DYN
            # set/get pair that is the pose read-back applied in _dynamics below.
            self._dynamics(rpm)
            self.last_clipped_action = np.array(rpm)      # keeps the action map's dtype (float32 on the THRUST path), BaseAviary.py:442
        self.rpy = bullet_euler_from_quaternion(self.quat)
        obs = self._computeObs()
        reward = self._computeReward()
        terminated = self._computeTerminated()
        truncated = self._computeTruncated()
        info = self._computeInfo()
        self.step_counter += self.PYB_STEPS_PER_CTRL
        if not terminated:
            self._update_state_post_step(action)
        return obs, reward, terminated, truncated, info

    def rescale_action(self, action):
        return rescale_action(action, self.physical_action_bounds)

    def _preprocessAction(self, action):
        if self.ACT_TYPE == ACT_THRUST:
            return thrust_to_rpm(action, self.physical_action_bounds, self.C)
        if self.ACT_TYPE == ACT_RPM:
            return rpm_action_to_rpm(action, self.C, self.numpy_legacy_cast)
        if self.ACT_TYPE == ACT_ONE_D_RPM:
            return np.repeat(rpm_action_to_rpm(np.asarray(action).reshape(-1)[:1], self.C, self.numpy_legacy_cast), 4)
        # ---- BaseSingleAgentAviary.py:180-223: the PID family; the controller sees the state vector of step entry
        state = self._getDroneStateVector()
        action = np.asarray(action).reshape(-1)     # dtype kept: VEL's unit vector / target velocity stay float32 for float32 actions
        kw = dict(control_timestep=self.CTRL_TIMESTEP, cur_pos=state[0:3], cur_quat=state[3:7],
                  cur_vel=state[10:13], cur_ang_vel=state[13:16])
        if self.ACT_TYPE == ACT_PID:
            next_pos = calculate_next_step(current_position=state[0:3], destination=action[0:3], step_size=1)
            rpm, _, _ = self.ctrl.computeControl(target_pos=next_pos, **kw)
            return rpm
        if self.ACT_TYPE == ACT_VEL:
            if np.linalg.norm(action[0:3]) != 0:
                v_unit_vector = action[0:3] / np.linalg.norm(action[0:3])
            else:
                v_unit_vector = np.zeros(3)
            rpm, _, _ = self.ctrl.computeControl(target_pos=state[0:3], target_rpy=np.array([0, 0, state[9]]),
                                                 target_vel=self.SPEED_LIMIT * np.abs(action[3]) * v_unit_vector, **kw)
            return rpm
        if self.ACT_TYPE == ACT_ONE_D_PID:
            rpm, _, _ = self.ctrl.computeControl(target_pos=state[0:3] + 0.1 * np.array([0, 0, action[0]]), **kw)
            return rpm
        raise ValueError(self.ACT_TYPE)

    # ---- BaseAviary._dynamics (BaseAviary.py:899-958) ----------------------
    def _dynamics(self, rpm):
        c = self.C
        dt = self.PYB_TIMESTEP
        pos, quat, vel, rpy_rates = self.pos, self.quat, self.vel, self.rpy_rates
        rotation = bullet_matrix_from_quaternion(quat)
        forces = np.array(rpm ** 2) * c.KF                       # float32 on the THRUST path
        if self.PHYSICS in (PHYSICS_DYN_GND, PHYSICS_DYN_GND_DRAG):
            forces = forces + self._ground_effect(rpm, rotation)   # extension (a7)
        thrust = np.array([0, 0, np.sum(forces)])
        thrust_world_frame = np.dot(rotation, thrust)
        force_world_frame = thrust_world_frame - np.array([0, 0, self.GRAVITY])
        if self.PHYSICS in (PHYSICS_DYN_DRAG, PHYSICS_DYN_GND_DRAG):
            force_world_frame = force_world_frame + self._drag(rotation)  # extension (a7)
        z_torques = np.array(rpm ** 2) * c.KM
        if self.DRONE_MODEL == MODEL_RACE:                          # :927-928
            z_torques = -z_torques
        z_torque = (-z_torques[0] + z_torques[1] - z_torques[2] + z_torques[3])
        if self.DRONE_MODEL in (MODEL_CF2X, MODEL_RACE):            # :930-932
            x_torque = (forces[0] + forces[1] - forces[2] - forces[3]) * (c.L / np.sqrt(2))
            y_torque = (-forces[0] + forces[1] + forces[2] - forces[3]) * (c.L / np.sqrt(2))
        else:                                                       # DroneModel.CF2P, :933-935
            x_torque = (forces[1] - forces[3]) * c.L
            y_torque = (-forces[0] + forces[2]) * c.L
        torques = np.array([x_torque, y_torque, z_torque], dtype=np.float64)
        torques = torques - np.cross(rpy_rates, np.dot(self.J, rpy_rates))
        rpy_rates_deriv = np.dot(self.J_INV, torques)
        no_pybullet_dyn_accs = force_world_frame / c.M
        vel = vel + dt * no_pybullet_dyn_accs
        rpy_rates = rpy_rates + dt * rpy_rates_deriv
        pos = pos + dt * vel
        quat = self._integrateQ(quat, rpy_rates, dt)             # TIMESTEP := PYB_TIMESTEP
        # set pose / velocity in Bullet, then read them back (:946-956, :596-598)
        self.pos = pos
        self.quat = bullet_pose_readback(quat)
        self.vel = vel
        self.ang_v = np.dot(rotation, rpy_rates)
        self.rpy_rates = rpy_rates

    @staticmethod
    def _integrateQ(quat, omega, dt):
        """BaseAviary.py:960-973."""
        omega_norm = np.linalg.norm(omega)
        p, q, r = omega
        if np.isclose(omega_norm, 0):
            return quat
        lambda_ = np.array([
            [0, r, -q, p],
            [-r, 0, p, q],
            [q, -p, 0, r],
            [-p, -q, -r, 0],
        ]) * .5
        theta = omega_norm * dt / 2
        return np.dot(np.eye(4) * np.cos(theta) + 2 / omega_norm * lambda_ * np.sin(theta), quat)

    # ---- extensions: same formulas as the PYB add-ons, applied inside DYN --
    def _drag(self, rotation):
        """BaseAviary.py:838-865: force R.(-DRAG_COEFF * v_world * sum(2 pi rpm_prev/60)),
        applied to link 4 in LINK_FRAME, i.e. Bullet rotates the vector by R once
        more -- reproduced literally."""
        return np.dot(rotation, self._drag_link(rotation))

    def _drag_link(self, rotation):
        """The forceObj BaseAviary._drag hands to p.applyExternalForce(..., flags=p.LINK_FRAME) (BaseAviary.py:855-865);
        pinned against the reference's own function by tests/golden/forces_ref.npz."""
        drag_factors = -1 * self.DRAG_COEFF * np.sum(np.array(2 * np.pi * self.last_clipped_action / 60))
        return np.dot(rotation, drag_factors * np.array(self.vel))

    def _ground_effect(self, rpm, rotation):
        """BaseAviary.py:798-834: per-prop extra thrust along body z."""
        c = self.C
        rpy = bullet_euler_from_quaternion(self.quat)
        offs = np.array([[x, y, 0.0] for (x, y) in c.PROP_XY])
        prop_heights = np.array([self.pos[2] + np.dot(rotation, offs[i])[2] for i in range(4)])
        prop_heights = np.clip(prop_heights, c.GND_EFF_H_CLIP, np.inf)
        gnd = (np.array(rpm, dtype=np.float64) ** 2 * c.KF * c.GND_EFF_COEFF
               * (c.PROP_RADIUS / (4 * prop_heights)) ** 2)
        if np.abs(rpy[0]) < np.pi / 2 and np.abs(rpy[1]) < np.pi / 2:
            return gnd
        return np.zeros(4)

    # ---- observation (PBDroneEnv.py:296-398) -------------------------------
    def _getDroneStateVector(self):
        return np.hstack([self.pos, self.quat, self.rpy, self.vel, self.ang_v,
                          self.last_clipped_action]).reshape(20, )

    def _clipAndNormalizeState(self, state):
        MAX_LIN_VEL_XY, MAX_LIN_VEL_Z = 3, 1
        MAX_PITCH_ROLL = np.pi
        clipped_rp = np.clip(state[7:9], -MAX_PITCH_ROLL, MAX_PITCH_ROLL)
        clipped_vel_xy = np.clip(state[10:12], -MAX_LIN_VEL_XY, MAX_LIN_VEL_XY)
        clipped_vel_z = np.clip(state[12], -MAX_LIN_VEL_Z, MAX_LIN_VEL_Z)
        normalized_pos_xy = state[0:2] / np.array([self._x_high, self._y_high])
        normalized_pos_z = state[2] / self._z_high
        normalized_rp = clipped_rp / MAX_PITCH_ROLL
        normalized_y = state[9] / np.pi
        normalized_vel_xy = clipped_vel_xy / MAX_LIN_VEL_XY
        normalized_vel_z = clipped_vel_z / MAX_LIN_VEL_XY       # sic (PBDroneEnv.py:382)
        n = np.linalg.norm(state[13:16])
        normalized_ang_vel = state[13:16] / n if n != 0 else state[13:16]
        return np.hstack([normalized_pos_xy, normalized_pos_z, state[3:7], normalized_rp,
                          normalized_y, normalized_vel_xy, normalized_vel_z,
                          normalized_ang_vel, state[16:20]]).reshape(20, )

    def _computeObs(self):
        obs = self._clipAndNormalizeState(self._getDroneStateVector())
        ret = np.hstack([obs[0:3], obs[7:10], obs[10:13], obs[13:16]]).reshape(12, )
        if self.include_distance:
            ret = np.append(ret, [self._distance_to_target / self._max_target_dist])
        ret = np.clip(ret, np.finfo(np.float32).min, np.finfo(np.float32).max)
        return ret.astype(np.float32)

    # ---- info / truncation / termination (PBDroneEnv.py:434-473) -----------
    def _computeInfo(self):
        return {"found_targets": self._current_target_index}

    def _computeTruncated(self):
        return bool(self._max_steps <= self._steps)

    def _computeTerminated(self):
        return bool(self._is_done or self._has_collision_occurred())

    def current_target(self):
        if self._current_target_index < len(self._target_points):
            return self._target_points[self._current_target_index]
        return None

    def _has_collision_occurred(self):
        """PBDroneEnv.py:678-707.  DYN never steps Bullet, so getContactPoints() is
        always empty; ``ground_contact`` is our documented analytic substitute
        (collision cylinder half-height vs the plane z = 0), off by default."""
        state = self.pos
        contact = bool(self.ground_contact and state[2] < self.C.COLLISION_H / 2)
        self.margins += [state[0] - self._x_high, state[0] - self._x_low, state[1] - self._y_high,
                         state[1] - self._y_low, state[2] - self._z_high]
        if self.ground_contact:
            self.margins.append(state[2] - self.C.COLLISION_H / 2)
        return bool(state[0] > self._x_high or state[0] < self._x_low or
                    state[1] > self._y_high or state[1] < self._y_low or
                    contact or
                    state[2] > self._z_high or
                    (self.cylinder and self.is_out_of_cylinder_bounds(state)))

    def is_out_of_cylinder_bounds(self, drone_position, circle_center=(0, 0, 1), extension_length=0.2):
        """PBDroneEnv.py:718-786 (worker-process numpy error state: 0/0 -> NaN ->
        comparison False)."""
        if self.circle:
            drone_vec = np.array(drone_position, dtype=np.float64)
            center_vec = np.array(circle_center, dtype=np.float64)
            center_to_drone_vec = drone_vec - center_vec
            center_to_drone_vec[2] = 0
            with np.errstate(all="ignore"):
                norm_vec = center_to_drone_vec / np.linalg.norm(center_to_drone_vec) * self.circle_radius
            closest_point = center_vec + norm_vec
            distance_from_closest_point = np.linalg.norm(drone_position - closest_point)
            self.margins.append(distance_from_closest_point - self._threshold)
            return bool(distance_from_closest_point > self._threshold)
        if self._current_target_index == 0:
            base1 = np.array(self.INIT_XYZS[0])
            base2 = np.array(self.current_target())
        else:
            base1 = np.array(self._target_points[self._current_target_index - 1])
            base2 = np.array(self.current_target())
        line_vec = base2 - base1
        line_length = np.linalg.norm(line_vec)
        if line_length == 0:
            self.margins.append(np.linalg.norm(drone_position - base1) - self._threshold)
            return bool(np.linalg.norm(drone_position - base1) > self._threshold)
        line_unit_vec = line_vec / line_length
        extended_point1 = base1 - extension_length * line_unit_vec
        extended_point2 = base2 + extension_length * line_unit_vec
        point1_to_drone_vec = drone_position - extended_point1
        projection_length = np.dot(point1_to_drone_vec, line_unit_vec)
        projection_length = np.clip(projection_length, 0, np.linalg.norm(extended_point2 - extended_point1))
        closest_point_on_line = extended_point1 + projection_length * line_unit_vec
        distance_from_line = np.linalg.norm(drone_position - closest_point_on_line)
        self.margins.append(distance_from_line - (self._threshold + extension_length))
        return bool(distance_from_line > self._threshold + extension_length)

    # ---- reward families other than the default ------------
    def _computeReward(self):
        rid = self.reward_id
        if rid == "default":
            return self._reward_waypoint(-10.0, 200, 75, 5, 3000, 3, (0.7, 0.3))
        if rid == "dummy":          # dummy_env.py:446-550, smoothness thresholds 0.1 / 0.1 (:587)
            return self._reward_waypoint(-10.0, 200, 75, 5, 3000, 3, (0.1, 0.1))
        if rid == "progress":       # PBDroneEnv reward with Rewarder.calculate_progress_reward x 2000 in place of the distance difference
            return self._reward_waypoint(-10.0, 200, 75, 5, 3000, 3, (0.7, 0.3), proj_w=2000)
        if rid == "thrustenv":      # ThrustEnv.py:368-463
            return self._reward_waypoint(-4.0, 1000, 25, 0, 20, 0, None)
        if rid == "her":
            return self._reward_her()
        if rid == "reaching":
            return self._reward_reaching()
        if rid in ("bootstrapped", "champ"):
            return self._reward_literature(rid)
        if rid == "hover":          # HoverAviary.py:65-76
            return -1 * np.linalg.norm(np.array([0, 0, 1]) - self.pos) ** 2
        if rid == "flythrugate":    # FlyThruGateAviary.py:100-112
            norm_ep_time = (self.step_counter / self.PYB_FREQ) / self.EPISODE_LEN_SEC
            return -10 * np.linalg.norm(np.array([0, -2 * norm_ep_time, 0.75]) - self.pos) ** 2
        raise ValueError(rid)

    def _reward_literature(self, rid):
        """Rewarder.BootstrappedImiVisionRewardCalculator.calculate_reward (Rewarder.py:66-104, arXiv 2403.12203) and
        Rewarder.ChampRewardCalculator.calculate_reward (Rewarder.py:107-150, Nature 2023).  The reference never calls
        these classes ("yet unused"); the wiring into PBDroneEnv's waypoint machine is ours and is the one the fixture
        generator uses around the reference's own calculator objects (tests/golden/make_ref_golden.py::_Literature):
        (prev_dis, dis) = the stale distance pair of the default reward, delta_cam = angle between the forward vector and
        the direction to the current target (0 on the target), a_t / a_t-1 = this step's action / _last_action,
        omega_t = rpy_rates, passed = target captured this step, crashed = collision before any capture."""
        crashed = bool(self._computeTerminated() and not self._is_done)
        passed = False
        if not crashed:
            self.margins.append(self._distance_to_target - self._threshold)
            if self._distance_to_target <= self._threshold:
                self._current_target_index += 1
                passed = True
                if self._current_target_index == len(self._target_points):
                    self._is_done = True
        target = self._target_points[min(self._current_target_index, len(self._target_points) - 1)]
        v = np.array(target, dtype=np.float64) - np.array(self.pos)
        n = np.linalg.norm(v)
        delta_cam = float(np.arccos(np.clip(np.dot(self.get_forward_vector(), v / n), -1.0, 1.0))) if n > 0 else 0.0
        a_t, a_tm1, omega = np.asarray(self._cur_action, np.float64), np.asarray(self._last_action, np.float64), self.rpy_rates
        prev_dis, dis = self._prev_distance_to_target, self._distance_to_target
        if rid == "bootstrapped":
            l1, l2, l3, l4, c1, c2 = 0.5, 0.025, 2e-4, 5e-4, 10, 4
            r = (l1 * (prev_dis - dis) + l2 * (l3 * (delta_cam ** 4)) - l3 * np.linalg.norm(a_t - a_tm1)
                 - l4 * np.linalg.norm(omega) + (c1 if passed else 0) + (-c2 if crashed else 0))
        else:
            l1, l2, l3, l4, l5, c1 = 1.0, 0.02, -10.0, -2e-4, -1e-4, 5.0
            r = (l1 * (prev_dis - dis) + l2 * np.exp(l3 * (delta_cam ** 4))
                 + (l4 * np.linalg.norm(omega) ** 2 + l5 * np.linalg.norm(a_t - a_tm1) ** 2)
                 - (c1 if self.pos[2] < 0 or crashed else 0))
        if not crashed:
            self._prev_distance_to_target = self._distance_to_target
        return r

    def _reward_her(self):
        """HerPBDroneEnv._computeReward (HerPBDroneEnv.py:314-398), first element of the returned tuple."""
        if self._computeTerminated() and not self._is_done:
            return -3000
        reward = 0.0
        distance_to_target = abs(np.linalg.norm(self._target_points[self._current_target_index] - self._current_position))
        reward += np.exp(-distance_to_target * 5) * 50
        reward += (self._prev_distance_to_target - distance_to_target) * 300
        self.margins.append(distance_to_target - self._threshold)
        if distance_to_target <= self._threshold:
            self._current_target_index += 1
            if self._current_target_index == len(self._target_points):
                reward += 1_000_000.0
                self._is_done = True
                return reward
            reward += 5000 * (self._discount ** (self._steps / 10))
            return reward
        self._prev_distance_to_target = distance_to_target
        return reward

    def _reward_reaching(self):
        """dummy_env.PBDroneEnv.progress_reward (dummy_env.py:617-643) == Rewarder.reaching_progress_reward
        (Rewarder.py:8-40)."""
        reward = 0
        dist_to_cent = np.linalg.norm(self._current_position - self._target_points[self._current_target_index])
        self.margins.append(dist_to_cent - self._threshold)
        if dist_to_cent <= self._threshold:
            self._current_target_index += 1
            reward += 3
        if self._current_target_index == len(self._target_points):
            self._is_done = True
            return 10
        dist_to_prev = np.linalg.norm(self._current_position - self._last_position)
        penalty_term = 0.01 * np.linalg.norm(self.pos.reshape(1, 3)[10:])    # empty slice of the (1, 3) array -> 0
        collision_penalty = -10.0 if self._has_collision_occurred() else 0.0
        reward += dist_to_prev - dist_to_cent - penalty_term + collision_penalty
        return reward

    # ---- reward (PBDroneEnv.py:475-607) and its constant variants -----------
    def calculate_progress_reward(self, pc_t, pc_t_minus_1, g1, g2):
        """Rewarder.py:43-62 / dummy_env.py:599-615."""
        def s(p):
            g_diff = g2 - g1
            return np.dot(p - g1, g_diff) / np.linalg.norm(g_diff) ** 2
        return s(pc_t) - s(pc_t_minus_1)

    def _reward_waypoint(self, crash, final, capture, capture_orient, progress_w, orient_w, smooth_thr, proj_w=0):
        if self._computeTerminated() and not self._is_done:
            return crash
        reward = np.float32(0.0)
        self.margins.append(self._distance_to_target - self._threshold)
        if self._distance_to_target <= self._threshold:
            self._current_target_index += 1
            if self._current_target_index == len(self._target_points):
                reward += final
                self._is_done = True
            else:
                reward += capture
                if capture_orient:
                    reward += self.orientation_reward(self.current_target()) * capture_orient
                self.just_found = True
        else:
            reward += (np.exp(-2 * self._distance_to_target)) * 3
            if proj_w:
                idx = self._current_target_index
                g1 = np.array(self.INIT_XYZS[0]) if idx == 0 else np.array(self._target_points[idx - 1])
                g2 = np.array(self._target_points[idx])
                prog = 0.0 if np.linalg.norm(g2 - g1) == 0 else proj_w * self.calculate_progress_reward(self.pos, self._pos_at_entry, g1, g2)
            else:
                prog = (self._prev_distance_to_target - self._distance_to_target) * progress_w
            reward += prog if not self.just_found else 0
            if orient_w:
                reward += self.orientation_reward(self.current_target()) * orient_w
            if smooth_thr is not None:
                reward += self.smoothness_reward(*smooth_thr)
            self.just_found = False
        self._prev_distance_to_target = self._distance_to_target
        return reward / 25

    def orientation_reward(self, target_pos):
        threshold_angle = np.radians(10)
        forward_vector = self.get_forward_vector()
        drone_to_target_vector = np.array(target_pos, dtype=np.float64) - np.array(self.pos)
        with np.errstate(all="ignore"):
            drone_to_target_vector = drone_to_target_vector / np.linalg.norm(drone_to_target_vector)
            angle = np.arccos(np.clip(np.dot(forward_vector, drone_to_target_vector), -1.0, 1.0))
        self.margins.append(angle - threshold_angle)
        return -1 if angle > threshold_angle else 0

    def get_forward_vector(self):
        euler = self.rpy
        return np.array([np.cos(euler[2]) * np.cos(euler[1]),
                         np.sin(euler[2]) * np.cos(euler[1]),
                         np.sin(euler[1])])

    def smoothness_reward(self, accel_threshold=0.7, ang_accel_threshold=0.3):
        lin_acc = np.linalg.norm(self.current_vel - self.prev_vel)
        ang_acc = np.linalg.norm(self.current_ang_v - self.prev_ang_v)
        self.margins += [lin_acc - accel_threshold, ang_acc - ang_accel_threshold]
        linear_penalty = -abs(lin_acc) if lin_acc > accel_threshold else 0
        angular_penalty = -abs(ang_acc) if ang_acc > ang_accel_threshold else 0
        return linear_penalty + angular_penalty

    # ---- post-step (PBDroneEnv.py:201-223) ----------------------------------
    def _update_state_post_step(self, action):
        self._steps += 1
        self._last_action = action                              # PBDroneEnv.py:205
        self._last_position = self._current_position.copy()     # dummy_env.py:196 (reaching-progress reward only)
        self._current_position = self.pos.copy()
        self.prev_vel, self.prev_ang_v = self.current_vel.copy(), self.current_ang_v.copy()
        self.current_vel, self.current_ang_v = self.vel.copy(), self.ang_v.copy()
        self._distance_to_target = np.linalg.norm(self.current_target() - self._current_position)

    # ---- state exchange with the CUDA path (tests upload identical states) --
    def set_state(self, st: dict):
        """Inverse of get_state (test helper): puts the env mid-episode in a given state."""
        f = lambda k: np.array(st[k], dtype=np.float64).copy()
        self.pos, self.quat, self.vel = f("pos"), f("quat"), f("vel")
        self.rpy_rates, self.ang_v = f("rpy_rates"), f("ang_v")
        self.rpy = bullet_euler_from_quaternion(self.quat)
        self.prev_vel, self.prev_ang_v = f("prev_vel"), f("prev_ang_v")
        self.current_vel, self.current_ang_v = self.vel.copy(), self.ang_v.copy()
        self._current_position = self.pos.copy()
        self._distance_to_target = float(st["dist"])
        self._prev_distance_to_target = float(st["prev_dist"])
        self._current_target_index = int(st["target_idx"])
        self._steps = int(st["steps"])
        self.just_found = bool(st["just_found"])
        self._is_done = False
        if "last_rpm" in st:
            self.last_clipped_action = f("last_rpm")

    def get_state(self) -> dict:
        return dict(pos=self.pos.copy(), quat=self.quat.copy(), vel=self.vel.copy(),
                    rpy_rates=self.rpy_rates.copy(), ang_v=self.ang_v.copy(),
                    prev_vel=self.prev_vel.copy(), prev_ang_v=self.prev_ang_v.copy(),
                    dist=float(self._distance_to_target), prev_dist=float(self._prev_distance_to_target),
                    target_idx=int(self._current_target_index), steps=int(self._steps),
                    just_found=bool(self.just_found), last_rpm=self.last_clipped_action.copy())


# --------------------------------------------------------------------------
# wrappers: normalize.py + SB3 Monitor + SubprocVecEnv worker auto-reset
# --------------------------------------------------------------------------


class OracleRunningMeanStd:
    """normalize.py:10-47."""

    def __init__(self, epsilon=1e-4, shape=()):
        self.mean = np.zeros(shape, "float64")
        self.var = np.ones(shape, "float64")
        self.count = epsilon

    def update(self, x):
        batch_mean = np.mean(x, axis=0)
        batch_var = np.var(x, axis=0)
        batch_count = x.shape[0]
        delta = batch_mean - self.mean
        tot_count = self.count + batch_count
        new_mean = self.mean + delta * batch_count / tot_count
        m_a = self.var * self.count
        m_b = batch_var * batch_count
        M2 = m_a + m_b + np.square(delta) * self.count * batch_count / tot_count
        self.mean, self.var, self.count = new_mean, M2 / tot_count, tot_count


class OracleWorker:
    """One SubprocVecEnv worker = Monitor(NormalizeObservation(PBDroneEnv)) with
    the worker loop's auto-reset (PBDroneSimulator.py:153-198; SB3 contract in
    SURVEY.md a17)."""

    def __init__(self, env: OracleDroneEnv, normalize_obs=True, norm_epsilon=1e-8,
                 normalize_reward=False, clip_reward=0.0, reward_gamma=0.99):
        self.env = env
        self.normalize_obs = normalize_obs
        # gym TransformReward(clip) then NormalizeReward, both inside Monitor (PBDroneSimulator.py:190-195);
        # NormalizeReward restated from normalize.py:100-147
        self.normalize_reward, self.clip_reward, self.reward_gamma = normalize_reward, clip_reward, reward_gamma
        self.return_rms = OracleRunningMeanStd(shape=())
        self.returns = np.zeros(1)
        self.obs_rms = OracleRunningMeanStd(shape=(env.obs_dim,))
        self.norm_epsilon = norm_epsilon
        self.ep_return = 0.0
        self.ep_len = 0
        self.last_terminated = self.last_truncated = False

    def _normalize(self, obs):
        if not self.normalize_obs:
            return obs
        self.obs_rms.update(np.array([obs]))
        return ((np.array([obs]) - self.obs_rms.mean) / np.sqrt(self.obs_rms.var + self.norm_epsilon))[0]

    def reset(self):
        obs, info = self.env.reset()
        self.ep_return, self.ep_len = 0.0, 0
        return self._normalize(obs), info

    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(action)
        self.last_terminated, self.last_truncated = bool(terminated), bool(truncated)
        self.last_step_ang_v_norm = float(np.linalg.norm(self.env.ang_v))   # test instrumentation
        obs = self._normalize(obs)
        if self.clip_reward > 0:
            reward = np.clip(reward, -self.clip_reward, self.clip_reward)
        if self.normalize_reward:
            rews = np.array([reward])
            self.returns = self.returns * self.reward_gamma + rews
            self.return_rms.update(self.returns)
            rews = rews / np.sqrt(self.return_rms.var + self.norm_epsilon)
            if terminated or truncated:
                self.returns[:] = 0.0
            reward = rews[0]
        self.ep_return += float(reward)
        self.ep_len += 1
        done = terminated or truncated
        info = dict(info)
        if done:
            info["episode"] = {"r": round(self.ep_return, 6), "l": self.ep_len}
        info["TimeLimit.truncated"] = truncated and not terminated
        if done:
            info["terminal_observation"] = obs
            obs, _ = self.reset()
        return obs, reward, done, info


# --------------------------------------------------------------------------
# tracks (Waypoints.py:108-139,172-197; PBDroneSimulator.py:89-105,129-130)
# --------------------------------------------------------------------------


def circle_track(radius=1, num_points=6, height=1, center=(0, 0, 0)):
    angles = np.linspace(0, 2 * np.pi, num_points + 1, endpoint=True)
    pts = np.zeros((num_points + 1, 3))
    pts[:, 0] = center[0] + radius * np.cos(angles)
    pts[:, 1] = center[1] + radius * np.sin(angles)
    pts[:, 2] = center[2] + height
    targets = list(pts)[1:]                       # circle=True pops the first (PBDroneSimulator.py:129-130)
    return np.array(targets), np.array([[radius, 0, center[2] + radius]], dtype=np.float64), np.array([-2, -2, 0, 2, 2, 2])


def reaching_track():
    arr = np.array([[-2.5, 4.5, 3], [10, 3.5, 1], [8, -4.5, 1], [-4.5, -6, 2],
                    [-5, -5, 2], [5, -1, 3], [2.5, 6, 3], [-2.5, 4.5, 3]], dtype=np.float64)
    for i in range(len(arr)):
        arr[i][2] += 3
        arr[i] /= 5
    return arr, np.array([arr[0]]), np.array([-4, -4, 0, 4, 4, 4])


def make_reference_env(track="circle", pyb_freq=240, ctrl_freq=240, max_steps=4096, **kw) -> OracleDroneEnv:
    """The env PBDroneSimulator.make_env builds (PBDroneSimulator.py:154-172)."""
    if track == "circle":
        targets, init, dim = circle_track()
        circle = True
    elif track == "reaching":
        targets, init, dim = reaching_track()
        circle = False
    else:
        raise ValueError(track)
    args = dict(target_points=targets, threshold=0.3, discount=0.999, max_steps=max_steps,
                aviary_dim=dim, initial_xyzs=init, pyb_freq=pyb_freq, ctrl_freq=ctrl_freq,
                act=ACT_THRUST, cylinder=True, circle=circle, include_distance=True,
                normalize_actions=True)
    args.update(kw)
    return OracleDroneEnv(**args)
