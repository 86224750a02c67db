"""drl_dronenavigation_b200 -- B200-native batched drone-navigation environment.

The hot path (one fused sm_100a kernel per control step) lives in ``csrc/`` behind the
C ABI of ``include/dronenav.h``; the modules here mirror the reference's Python
interface for that path.  Importing this package never imports ``oracle/``.
"""
from .enums import ActionType, DroneModel, ImageType, ObservationType, Physics  # noqa: F401
from . import waypoints as Waypoints  # noqa: F401
from .waypoints import Track, dilate_targets, track_targets  # noqa: F401
from .constants import CF2X, DroneParameters, parse_urdf_parameters  # noqa: F401

__all__ = ["ActionType", "DroneModel", "ImageType", "ObservationType", "Physics", "Waypoints", "Track",
           "dilate_targets", "track_targets", "CF2X", "DroneParameters", "parse_urdf_parameters",
           "BatchedDroneEnv", "GpuDroneVecEnv", "PBDroneEnv"]


def __getattr__(name):   # torch / CUDA-dependent classes are imported on first use
    if name == "BatchedDroneEnv":
        from .batched_env import BatchedDroneEnv
        return BatchedDroneEnv
    if name == "GpuDroneVecEnv":
        from .vec_env import GpuDroneVecEnv
        return GpuDroneVecEnv
    if name == "PBDroneEnv":
        from .env import PBDroneEnv
        return PBDroneEnv
    raise AttributeError(name)
