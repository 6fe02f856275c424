"""ctypes wrapper of oracle/libcavgym_oracle.so (TEST INFRASTRUCTURE ONLY).

Mirrors the engine's API on numpy arrays with the engine's SoA layout
([M][4][N] state, [M][2][N] actions, ...) so tests compare arrays directly.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from cavgym_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libcavgym_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "cavgym_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "cavgym.h")
    if force or not os.path.isfile(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off", "-fno-fast-math",
               "-o", _LIB_PATH, src, "-lm", "-lpthread"]
        subprocess.run(cmd, check=True, cwd=_HERE)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, i64, f64p = C.c_void_p, C.c_int64, C.c_void_p
        L.cav_oracle_create.restype = C.c_int
        L.cav_oracle_create.argtypes = [C.POINTER(_abi.CavScenario), i64, C.c_uint64, C.POINTER(vp)]
        L.cav_oracle_destroy.argtypes = [vp]
        L.cav_oracle_set_shard.argtypes = [vp, i64]
        L.cav_oracle_set_threads.argtypes = [vp, C.c_int]
        L.cav_oracle_set_tangent_tolerance.argtypes = [vp, C.c_double]
        L.cav_oracle_set_global_timestep.argtypes = [vp, i64]
        L.cav_oracle_set_uniform_override.argtypes = [vp, f64p]
        L.cav_oracle_set_spawn_override.argtypes = [vp, f64p]
        L.cav_oracle_reset.argtypes = [vp, vp, f64p]
        L.cav_oracle_step.argtypes = [vp, f64p, f64p, f64p, vp, vp, vp]
        L.cav_oracle_rollout.argtypes = [vp, C.c_int, C.c_int]
        L.cav_oracle_replay.argtypes = [vp, C.c_int, f64p, f64p, f64p, vp, vp, vp]
        L.cav_oracle_stats.argtypes = [vp, C.POINTER(C.c_int64)]
        L.cav_oracle_info.argtypes = [vp, f64p, f64p]
        for name in ("state", "action", "agent_state", "liveness", "timestep", "winner", "done", "error"):
            fn = getattr(L, f"cav_oracle_{name}_ptr")
            fn.restype, fn.argtypes = vp, [vp]
        L.cav_oracle_philox.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.cav_oracle_spawn.argtypes = [C.POINTER(_abi.CavSpawn), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.cav_oracle_make_box.argtypes = [C.c_double] * 5 + [C.POINTER(_abi.CavQuad)]
        L.cav_oracle_intersects.restype = C.c_int
        L.cav_oracle_intersects.argtypes = [C.POINTER(_abi.CavQuad)] * 2
        L.cav_oracle_contains.restype = C.c_int
        L.cav_oracle_contains.argtypes = [C.POINTER(_abi.CavQuad)] * 2
        L.cav_oracle_percentage_intersects.restype = C.c_double
        L.cav_oracle_percentage_intersects.argtypes = [C.POINTER(_abi.CavQuad)] * 2
        L.cav_oracle_stopping_zones.restype = C.c_int
        L.cav_oracle_stopping_zones.argtypes = [C.POINTER(_abi.CavBodyType), C.POINTER(C.c_double), C.c_double,
                                                C.POINTER(_abi.CavQuad), C.POINTER(_abi.CavQuad)]
        L.cav_oracle_dynamic_body_step.argtypes = [C.POINTER(_abi.CavBodyType), C.POINTER(C.c_double), C.c_double, C.c_double, C.c_double]
        L.cav_oracle_make_steering_action.restype = C.c_double
        L.cav_oracle_make_steering_action.argtypes = [C.POINTER(_abi.CavBodyType), C.POINTER(C.c_double), C.c_double, C.c_double]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def quad(points):
    q = _abi.CavQuad()
    for i, (x, y) in enumerate(points):
        q.x[i], q.y[i] = float(x), float(y)
    return q


def quad_points(q):
    return [(q.x[i], q.y[i]) for i in range(4)]


class Oracle:
    """CPU engine with the semantics of cavgym_create/reset/step/rollout, fp64 only."""

    def __init__(self, compiled, n_envs, seed=0, threads=1):
        self.compiled = compiled
        self.n, self.m = int(n_envs), compiled.n_bodies
        self._h = C.c_void_p()
        rc = lib().cav_oracle_create(compiled.pointer(), self.n, seed, C.byref(self._h))
        if rc != 0:
            raise RuntimeError(f"cav_oracle_create failed: {rc}")
        lib().cav_oracle_set_threads(self._h, threads)
        self._keep = {}

    def close(self):
        if self._h:
            lib().cav_oracle_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        self.close()

    def _view(self, name, shape, dtype):
        ptr = getattr(lib(), f"cav_oracle_{name}_ptr")(self._h)
        count = int(np.prod(shape))
        ctype = np.ctypeslib.as_ctypes_type(np.dtype(dtype))
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(count,)).reshape(shape)

    @property
    def state(self):
        return self._view("state", (self.m, 4, self.n), np.float64)

    @property
    def actions_taken(self):
        return self._view("action", (self.m, 2, self.n), np.float64)

    @property
    def agent_state(self):
        return self._view("agent_state", (self.m, _abi.CAV_AGENT_WORDS, self.n), np.float64)

    @property
    def liveness(self):
        return self._view("liveness", (self.m, self.n), np.int32)

    @property
    def timestep(self):
        return self._view("timestep", (self.n,), np.int32)

    @property
    def done_latch(self):
        return self._view("done", (self.n,), np.uint8)

    @property
    def winner_latch(self):
        return self._view("winner", (self.n,), np.int32)

    @property
    def error(self):
        return self._view("error", (self.n,), np.uint8)

    def set_shard(self, offset):
        lib().cav_oracle_set_shard(self._h, int(offset))

    def set_threads(self, threads):
        lib().cav_oracle_set_threads(self._h, int(threads))

    def set_tangent_tolerance(self, tau):
        lib().cav_oracle_set_tangent_tolerance(self._h, float(tau))

    def set_global_timestep(self, t):
        lib().cav_oracle_set_global_timestep(self._h, int(t))

    def set_uniform_override(self, u):
        self._keep["uni"] = None if u is None else np.ascontiguousarray(u, dtype=np.float64)
        lib().cav_oracle_set_uniform_override(self._h, _ptr(self._keep["uni"]))

    def set_spawn_override(self, u):
        self._keep["spawn"] = None if u is None else np.ascontiguousarray(u, dtype=np.float64)
        lib().cav_oracle_set_spawn_override(self._h, _ptr(self._keep["spawn"]))

    def reset(self, mask=None, init_state=None):
        mask = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        init_state = None if init_state is None else np.ascontiguousarray(init_state, dtype=np.float64)
        lib().cav_oracle_reset(self._h, _ptr(mask), _ptr(init_state))

    def step(self, actions=None):
        n, m = self.n, self.m
        actions = None if actions is None else np.ascontiguousarray(actions, dtype=np.float64)
        out = (np.empty((m, 4, n)), np.empty((m, n)), np.empty(n, np.uint8), np.empty(n, np.int32), np.empty(n, np.uint8))
        lib().cav_oracle_step(self._h, _ptr(actions), *[_ptr(a) for a in out])
        return out

    def rollout(self, n_steps, auto_reset=True):
        lib().cav_oracle_rollout(self._h, int(n_steps), int(bool(auto_reset)))

    def replay(self, actions, outputs=True):
        actions = np.ascontiguousarray(actions, dtype=np.float64)
        t, m, n = actions.shape[0], self.m, self.n
        if not outputs:
            lib().cav_oracle_replay(self._h, t, _ptr(actions), None, None, None, None, None)
            return None
        out = (np.empty((t, m, 4, n)), np.empty((t, m, n)), np.empty((t, n), np.uint8), np.empty((t, n), np.int32),
               np.empty((t, n), np.uint8))
        lib().cav_oracle_replay(self._h, t, _ptr(actions), *[_ptr(a) for a in out])
        return out

    def info(self):
        """CAVEnv.info for the current state: (body_polygons [M, 8, N], road_angles [M, N], NaN = None)."""
        polygons, angles = np.empty((self.m, 8, self.n)), np.empty((self.m, self.n))
        lib().cav_oracle_info(self._h, _ptr(polygons), _ptr(angles))
        return polygons, angles

    def stats(self):
        out = (C.c_int64 * _abi.CAV_N_STATS)()
        lib().cav_oracle_stats(self._h, out)
        return dict(zip(_abi.STAT_NAMES, [int(v) for v in out]))


def philox(counter, key):
    ctr = (C.c_uint32 * 4)(*counter)
    k = (C.c_uint32 * 2)(*key)
    out = (C.c_uint32 * 4)()
    lib().cav_oracle_philox(ctr, k, out)
    return [int(v) for v in out]
