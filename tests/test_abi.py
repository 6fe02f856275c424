"""The C-ABI shared library loads and exports every symbol include/cavgym.h declares (no compute, CPU-only)."""
import ctypes
import os
import re

from cavgym_b200 import _abi, _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "cavgym.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cavgym_\w+)\s*\(", text)))


def test_header_and_ctypes_mirror_agree():
    assert declared_symbols() == sorted(_abi.PROTOTYPES)


def test_library_exports_every_declared_symbol():
    lib = _native.load()
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.cavgym_version()


def test_struct_layout_matches_header_sizes():
    # sizes computed from the C declarations: doubles are 8-byte aligned, no packing pragmas
    assert ctypes.sizeof(_abi.CavQuad) == 64
    assert ctypes.sizeof(_abi.CavBodyType) == 72
    assert ctypes.sizeof(_abi.CavSpawn) == 8 + 2 * 64 + 4 * 8 + 8
    assert ctypes.sizeof(_abi.CavBody) == 24 + 16 + 32 + 64
    assert ctypes.sizeof(_abi.CavScenario) == 32 + 8 + 7 * 8 + 32 + 4 * 64 + 8 * 64 + 8 * 72 + 16


def test_bad_arguments_return_codes_without_a_gpu():
    lib = _native.load()
    assert lib.cavgym_create(None, 1, 0, 0, 0, None) == -22
    assert b"NULL" in lib.cavgym_last_error()
    assert lib.cavgym_destroy(None) == 0
    assert lib.cavgym_set_global_timestep(None, 0) == -22
    assert lib.cavgym_info(None, None, None, None) == -22 and lib.cavgym_step(None, None, None, None, None, None, None, None) == -22
    assert lib.cavgym_set_dense_path(None, 1) == -22 and lib.cavgym_rollout(None, 1, 1, None) == -22
