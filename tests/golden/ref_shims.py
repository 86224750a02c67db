"""Import shims that let the UNMODIFIED reference (``/root/reference``) be imported and stepped in this
container, where ``pybullet``, ``gymnasium``, ``gym``, ``stable_baselines3`` and ``matplotlib`` are absent.

TEST INFRASTRUCTURE ONLY: used by ``tests/golden/make_ref_golden.py`` to mint fixtures from the reference's own
code.  Nothing here is imported by the product package, by ``-m gpu`` tests, ``smoke()`` or ``bench.py``
(``/root/reference`` does not exist on the GPU box; only the committed ``.npz`` fixtures travel).

What is shimmed and how faithfully:

* ``pybullet``  -- a rigid-body *state store* (no dynamics, no contacts) with Bullet's published maths for the
  five functions the DYN path depends on.  Restated from the bullet3 sources, double precision as in the
  pybullet wheel (``BT_USE_DOUBLE_PRECISION``):
    - ``getQuaternionFromEuler``            btQuaternion::setEulerZYX(yaw, pitch, roll) then normalize (pybullet.c)
    - ``getMatrixFromQuaternion``           btMatrix3x3::setRotation  (s = 2/|q|^2)
    - ``getEulerFromQuaternion``            pybullet.c (sarg = -2(xz - wy), gimbal branches at |sarg| >= 0.99999)
    - ``resetBasePositionAndOrientation`` / ``getBasePositionAndOrientation``
                                            btMultiBody keeps the quaternion; the read-back goes through
                                            btTransform(q) -> btMatrix3x3::getRotation, i.e. a UNIT quaternion
                                            (possibly sign-flipped) of the same rotation
    - ``resetBaseVelocity`` / ``getBaseVelocity``  stored as given
    - ``getContactPoints`` -> ()  (DYN never calls stepSimulation, so Bullet has no contacts to report)
  Every other attribute is a no-op returning 0 (GUI, cameras, debug items, gravity, time step ...).
* ``gymnasium`` / ``gym`` -- ``Env``, ``Wrapper``/``core.Wrapper``, ``spaces.Box`` with just the attributes used.
* ``gymnasium.envs.registration.register`` -- no-op (the vendored upstream package registers env ids when
  ``BaseControl._getURDFParameter`` imports it through ``pkg_resources`` to find ``assets/cf2x.urdf``).
* ``stable_baselines3.common.running_mean_std`` -- imported by PBDroneEnv.py:21 but unused on the path.
* ``matplotlib``, ``mpl_toolkits``, ``torchviz``, ``graphviz``, ``hiddenlayer``, ``pybullet_data`` -- inert.
"""
from __future__ import annotations

import math
import sys
import types

import numpy as np


class _Inert(types.ModuleType):
    """Module whose every attribute is an inert callable / sub-module (plotting & GUI packages)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = _Inert(self.__name__ + "." + name)
        sys.modules.setdefault(sub.__name__, sub)
        setattr(self, name, sub)
        return sub

    def __call__(self, *a, **k):
        return _Inert(self.__name__ + "()")

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)


# --------------------------------------------------------------------------------------------
# pybullet state store
# --------------------------------------------------------------------------------------------
def _bt_matrix_from_quat(q):
    x, y, z, w = (float(v) for v in q)
    d = x * x + y * y + z * z + w * w
    s = 2.0 / d
    xs, ys, zs = x * s, y * s, z * s
    wx, wy, wz = w * xs, w * ys, w * zs
    xx, xy, xz = x * xs, x * ys, x * zs
    yy, yz, zz = y * ys, y * zs, z * zs
    return ((1.0 - (yy + zz), xy - wz, xz + wy),
            (xy + wz, 1.0 - (xx + zz), yz - wx),
            (xz - wy, yz + wx, 1.0 - (xx + yy)))


def _bt_quat_from_matrix(m):
    """btMatrix3x3::getRotation."""
    trace = m[0][0] + m[1][1] + m[2][2]
    t = [0.0, 0.0, 0.0, 0.0]
    if trace > 0.0:
        s = math.sqrt(trace + 1.0)
        t[3] = s * 0.5
        s = 0.5 / s
        t[0] = (m[2][1] - m[1][2]) * s
        t[1] = (m[0][2] - m[2][0]) * s
        t[2] = (m[1][0] - m[0][1]) * s
    else:
        i = 0 if m[0][0] >= m[1][1] else 1
        if m[2][2] > m[i][i]:
            i = 2
        j, k = (i + 1) % 3, (i + 2) % 3
        s = math.sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0)
        t[i] = s * 0.5
        s = 0.5 / s
        t[3] = (m[k][j] - m[j][k]) * s
        t[j] = (m[j][i] + m[i][j]) * s
        t[k] = (m[k][i] + m[i][k]) * s
    return tuple(t)


class _PyBullet(types.ModuleType):
    DIRECT, GUI = 2, 1
    LINK_FRAME, WORLD_FRAME = 1, 2
    URDF_USE_INERTIA_FROM_FILE = 2
    GEOM_CYLINDER, GEOM_SPHERE, GEOM_BOX = 4, 2, 3

    def __init__(self):
        super().__init__("pybullet")
        self._worlds = {}       # physicsClientId -> {"bodies": {uid: state}, "next": int}; one world per p.connect()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        if name.isupper():
            return 0
        return lambda *a, **k: 0

    # ---- world -----------------------------------------------------------------------------
    def connect(self, *a, **k):
        cid = len(self._worlds)
        self._worlds[cid] = {"bodies": {}, "next": 0}
        return cid

    def _body(self, uid, k):
        return self._worlds[k.get("physicsClientId", 0)]["bodies"][int(uid)]

    def resetSimulation(self, *a, **k):
        w = self._worlds[k.get("physicsClientId", 0)]
        w["bodies"].clear()
        w["next"] = 0

    def loadURDF(self, fileName, basePosition=(0.0, 0.0, 0.0), baseOrientation=(0.0, 0.0, 0.0, 1.0), *a, **k):
        w = self._worlds[k.get("physicsClientId", 0)]
        uid = w["next"]
        w["next"] += 1
        w["bodies"][uid] = {"pos": tuple(float(v) for v in basePosition),
                            "quat": tuple(float(v) for v in baseOrientation),
                            "lin": (0.0, 0.0, 0.0), "ang": (0.0, 0.0, 0.0)}
        return uid

    def stepSimulation(self, *a, **k):
        raise RuntimeError("the DYN path never calls p.stepSimulation (BaseAviary.py:439-440)")

    def getContactPoints(self, *a, **k):
        return ()

    # ---- maths -----------------------------------------------------------------------------
    def getQuaternionFromEuler(self, rpy, *a, **k):
        roll, pitch, yaw = (float(v) for v in rpy)
        hy, hp, hr = yaw * 0.5, pitch * 0.5, roll * 0.5
        cy, sy, cp, sp, cr, sr = math.cos(hy), math.sin(hy), math.cos(hp), math.sin(hp), math.cos(hr), math.sin(hr)
        q = (sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy, cr * cp * cy + sr * sp * sy)
        n = math.sqrt(sum(v * v for v in q))
        return tuple(v / n for v in q)

    def getMatrixFromQuaternion(self, q, *a, **k):
        m = _bt_matrix_from_quat(q)
        return m[0] + m[1] + m[2]

    def getEulerFromQuaternion(self, q, *a, **k):
        x, y, z, w = (float(v) for v in q)
        sqx, sqy, sqz, squ = x * x, y * y, z * z, w * w
        sarg = -2.0 * (x * z - w * y)
        if sarg <= -0.99999:
            return (0.0, -0.5 * math.pi, 2.0 * math.atan2(x, -y))
        if sarg >= 0.99999:
            return (0.0, 0.5 * math.pi, 2.0 * math.atan2(-x, y))
        return (math.atan2(2.0 * (y * z + w * x), squ - sqx - sqy + sqz), math.asin(sarg),
                math.atan2(2.0 * (x * y + w * z), squ + sqx - sqy - sqz))

    # ---- body state --------------------------------------------------------------------------
    def resetBasePositionAndOrientation(self, uid, pos, orn, *a, **k):
        b = self._body(uid, k)
        b["pos"] = tuple(float(v) for v in pos)
        b["quat"] = tuple(float(v) for v in orn)

    def getBasePositionAndOrientation(self, uid, *a, **k):
        b = self._body(uid, k)
        return b["pos"], _bt_quat_from_matrix(_bt_matrix_from_quat(b["quat"]))

    def resetBaseVelocity(self, uid, linearVelocity=None, angularVelocity=None, *a, **k):
        b = self._body(uid, k)
        if linearVelocity is not None:
            b["lin"] = tuple(float(v) for v in linearVelocity)
        if angularVelocity is not None:
            b["ang"] = tuple(float(v) for v in angularVelocity)

    def getBaseVelocity(self, uid, *a, **k):
        b = self._body(uid, k)
        return b["lin"], b["ang"]

    # ---- external forces: recorded, never integrated (the PYB_* force models of BaseAviary.py:798-895 hand their result
    #      to Bullet; the recorder lets a fixture pin the numbers the reference computes) ----------------------------------
    force_log = None            # list while recording: (kind, body uid, link index, vector, posObj, flags)

    def applyExternalForce(self, objectUniqueId, linkIndex, forceObj, posObj, flags, *a, **k):
        if self.force_log is not None:
            self.force_log.append(("force", int(objectUniqueId), int(linkIndex), tuple(float(v) for v in forceObj),
                                   tuple(float(v) for v in posObj), int(flags)))
        return 0

    def applyExternalTorque(self, objectUniqueId, linkIndex, torqueObj, flags, *a, **k):
        if self.force_log is not None:
            self.force_log.append(("torque", int(objectUniqueId), int(linkIndex), tuple(float(v) for v in torqueObj), (), int(flags)))
        return 0

    # propeller links 0..3 and the centre-of-mass link 4 of cf2x.urdf are fixed joints: world position = base position +
    # R(base quaternion) . joint origin (Bullet forward kinematics for fixed joints; joint origins read from the reference's URDF
    # by the caller into `link_offsets`)
    link_offsets = None

    def getLinkStates(self, bodyUniqueId, linkIndices, *a, **k):
        b = self._body(bodyUniqueId, k)
        m = _bt_matrix_from_quat(b["quat"])
        out = []
        for li in linkIndices:
            off = self.link_offsets[int(li)]
            world = tuple(b["pos"][r] + sum(m[r][c] * off[c] for c in range(3)) for r in range(3))
            out.append((world, b["quat"], (0.0, 0.0, 0.0), (0.0, 0.0, 0.0, 1.0), world, b["quat"], b["lin"], b["ang"]))
        return tuple(out)


# --------------------------------------------------------------------------------------------
# gymnasium / gym
# --------------------------------------------------------------------------------------------
class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        self.dtype = np.dtype(dtype)
        if shape is None:
            shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
        self.shape = tuple(shape)
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()
        self._rng = np.random.default_rng(seed)

    def sample(self):
        return self._rng.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low}, {self.high}, {self.shape}, {self.dtype})"


class Env:
    metadata = {}
    np_random = None

    def reset(self, seed=None, options=None):
        if seed is not None or self.np_random is None:
            self.np_random = np.random.default_rng(seed)

    def close(self):
        pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    def step(self, action):
        return self.env.step(action)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)


def _gym_module(name):
    m = types.ModuleType(name)
    m.Env, m.Wrapper = Env, Wrapper
    spaces = types.ModuleType(name + ".spaces")
    spaces.Box = Box
    spaces.Dict = dict
    spaces.Space = object
    spaces.__path__ = []                       # lets "from gymnasium.spaces.space import Space" resolve
    space = types.ModuleType(name + ".spaces.space")
    space.Space = object
    spaces.space = space
    sys.modules[name + ".spaces.space"] = space
    core = types.ModuleType(name + ".core")
    core.Wrapper, core.Env = Wrapper, Env
    m.spaces, m.core = spaces, core
    sys.modules[name + ".spaces"] = spaces
    sys.modules[name + ".core"] = core
    return m


def install(reference_root="/root/reference"):
    """Registers the shims in ``sys.modules`` and puts the reference on ``sys.path``.  Idempotent."""
    if "pybullet" not in sys.modules or not isinstance(sys.modules["pybullet"], _PyBullet):
        sys.modules["pybullet"] = _PyBullet()
    for name in ("gymnasium", "gym"):
        if name not in sys.modules:
            sys.modules[name] = _gym_module(name)
        if name + ".envs.registration" not in sys.modules:
            # the vendored gym_pybullet_drones/__init__.py registers its env ids on import (BaseControl._getURDFParameter
            # imports that package through pkg_resources to locate assets/cf2x.urdf): registration is a no-op here
            g = sys.modules[name]
            if not hasattr(g, "__path__"):
                g.__path__ = []
            envs, reg = types.ModuleType(name + ".envs"), types.ModuleType(name + ".envs.registration")
            envs.__path__ = []
            reg.register = lambda *a, **k: None
            g.envs, envs.registration = envs, reg
            sys.modules[name + ".envs"], sys.modules[name + ".envs.registration"] = envs, reg
    for name in ("pybullet_data", "matplotlib", "matplotlib.pyplot", "matplotlib.collections", "matplotlib.animation",
                 "mpl_toolkits", "mpl_toolkits.mplot3d", "torchviz", "graphviz", "hiddenlayer"):
        if name not in sys.modules:
            sys.modules[name] = _Inert(name)
    if "stable_baselines3" not in sys.modules:
        sb3 = types.ModuleType("stable_baselines3")
        common = types.ModuleType("stable_baselines3.common")
        rms = types.ModuleType("stable_baselines3.common.running_mean_std")
        rms.RunningMeanStd = type("RunningMeanStd", (), {"__init__": lambda self, *a, **k: None})
        sb3.common, common.running_mean_std = common, rms
        sys.modules.update({"stable_baselines3": sb3, "stable_baselines3.common": common,
                            "stable_baselines3.common.running_mean_std": rms})
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    return sys.modules["pybullet"]
