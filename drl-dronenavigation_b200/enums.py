"""Enumerations of the reference's API surface (Sol/PyBullet/enums.py:3-50).

Values are the reference's strings so saved configs / CLI flags keep working.  Only the
members marked ON PATH are accepted by the GPU environment; the others exist so that
reference code which names them still imports, and fail loudly when selected.
"""
from enum import Enum


class DroneModel(Enum):
    CF2X = "cf2x"    # ON PATH (the only model the reference itself can construct)
    CF2P = "cf2p"    # ON PATH: BaseAviary._dynamics' + frame branch (RPM and PID action types)
    RACE = "racer"   # ON PATH: X frame with reversed propeller spin (RPM action types)


class Physics(Enum):
    PYB = "pyb"                            # the CUDA env always integrates the DYN model
    DYN = "dyn"                            # ON PATH
    PYB_GND = "pyb_gnd"                    # ON PATH as DYN + ground effect
    PYB_DRAG = "pyb_drag"                  # ON PATH as DYN + drag
    PYB_DW = "pyb_dw"                      # multi-drone only
    PYB_GND_DRAG_DW = "pyb_gnd_drag_dw"    # ON PATH as DYN + ground effect + drag (downwash needs >1 drone)


class ImageType(Enum):
    RGB = 0
    DEP = 1
    SEG = 2
    BW = 3


class ActionType(Enum):
    RPM = "rpm"                # ON PATH
    PID = "pid"                # ON PATH (DSLPIDControl fused into the step kernel)
    VEL = "vel"                # ON PATH
    ONE_D_RPM = "one_d_rpm"    # ON PATH
    ONE_D_PID = "one_d_pid"    # ON PATH
    THRUST = "thrust"          # ON PATH (what PBDroneSimulator.make_env selects)


class ObservationType(Enum):
    KIN = "kin"                # ON PATH
    RGB = "rgb"
