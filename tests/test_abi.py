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


def test_create_validates_the_scenario_tables_without_a_gpu():
    """cavgym_create checks everything the kernels index fixed arrays with BEFORE it touches the device: spawn box /
    orientation counts (spawn_body), spawn and type ids, positive time resolution / viewer width / max_timesteps."""
    import copy
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import compile_from_meta, load_golden
    lib = _native.load()
    meta, _ = load_golden("pedestrians_rc_seed0")
    handle = ctypes.c_void_p()

    def create(mutate):
        compiled = compile_from_meta(copy.deepcopy(meta))
        mutate(compiled.struct)
        return lib.cavgym_create(compiled.pointer(), 4, 0, 0, 0, ctypes.byref(handle)), lib.cavgym_last_error()

    def set_spawn(field, value):
        def mutate(sc):
            setattr(sc.spawns[0], field, value)
        return mutate

    for field, value in (("n_boxes", 0), ("n_boxes", _abi.CAV_MAX_SPAWN_BOXES + 1), ("n_orientations", 0),
                         ("n_orientations", _abi.CAV_MAX_SPAWN_ORIENT + 1)):
        code, message = create(set_spawn(field, value))
        assert code == -22 and b"spawn" in message, (field, value)
    for field, value, word in (("n_spawns", -1, b"n_spawns"), ("time_resolution", 0.0, b"time_resolution"),
                               ("viewer_width", -1.0, b"viewer_width"), ("max_timesteps", 0, b"max_timesteps"), ("n_roads", 0, b"road")):
        code, message = create(lambda sc, f=field, v=value: setattr(sc, f, v))
        assert code == -22 and word in message, field
    assert handle.value is None


def test_host_and_episode_entry_points_refuse_a_null_engine():
    lib = _native.load()
    n_rows, dropped = ctypes.c_int64(), ctypes.c_int64()
    assert lib.cavgym_reset_host(None, None, None, None) == -22
    assert lib.cavgym_step_host(None, None, None, None, None, None, None) == -22
    assert lib.cavgym_set_episode_log(None, 16) == -22
    assert lib.cavgym_drain_episodes(None, None, 0, ctypes.byref(n_rows), ctypes.byref(dropped)) == -22
    assert ctypes.sizeof(_abi.CavEpisodeRow) == 24
