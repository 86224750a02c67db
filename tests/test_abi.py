"""The C-ABI library builds, loads and exports every symbol include/dronenav.h declares
(no compute calls: this file runs without a GPU)."""
import ctypes as C
import os
import re

from tests.conftest import ROOT


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "dronenav.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dn_[a-z_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    names = _declared_functions()
    for must in ("dn_create", "dn_destroy", "dn_reset", "dn_step", "dn_step_many", "dn_get_state",
                 "dn_set_state", "dn_episode_stats", "dn_last_error", "dn_abi_version"):
        assert must in names


def test_library_exports_every_declared_symbol(built_lib):
    from drl_dronenavigation_b200 import _lib
    names = _declared_functions()
    assert set(names) == set(_lib.SYMBOLS), "ctypes table and header disagree"
    for n in names:
        assert hasattr(built_lib, n), f"{n} not exported by libdronenav.so"
    assert built_lib.dn_abi_version() == _lib.DN_ABI_VERSION


def test_struct_layouts_match_header(built_lib):
    from drl_dronenavigation_b200 import _lib
    # sizes implied by include/dronenav.h on LP64
    assert C.sizeof(_lib.dn_step_io) == 8 * 8
    assert C.sizeof(_lib.dn_state_view) == 20 * 8
    assert C.sizeof(_lib.dn_stats) == 7 * 8
    assert C.sizeof(_lib.dn_config) == 4 + 4 + 8 + 8 + 12 * 4 + 2 * 8 + 12 * 8 + 4 + 4 + 2 * 8 + 8


def test_bad_config_is_rejected_without_a_gpu(built_lib):
    from drl_dronenavigation_b200 import _lib
    h = C.c_void_p()
    cfg = _lib.dn_config()
    cfg.abi_version = 999
    assert built_lib.dn_create(C.byref(cfg), 0, C.byref(h)) == -1
    assert b"abi_version" in built_lib.dn_last_error()
    cfg.abi_version = _lib.DN_ABI_VERSION
    cfg.num_envs, cfg.pyb_freq, cfg.ctrl_freq = 4, 240, 7
    assert built_lib.dn_create(C.byref(cfg), 0, C.byref(h)) == -1
    assert b"divisible" in built_lib.dn_last_error()


def test_no_cpu_fallback_in_product_code():
    """The package never imports oracle/ (the oracle is test infrastructure only)."""
    pkg = os.path.join(ROOT, "drl-dronenavigation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
