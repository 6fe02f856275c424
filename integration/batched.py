"""library/batched.py — the module a CAV-Gym maintainer adds to the reference tree to put `libcavgym_sm100.so` behind
the `CAVEnv` protocol.  Complete and runnable: ctypes against the C-ABI of include/cavgym.h, numpy
host arrays, no torch.  It takes the objects the reference already builds —

    np_seed, env, agents, keyboard_agent = config.setup()          # reference config.py:272-415
    batch = BatchedCAVEnv(env, num_envs=65536, seed=np_seed)
    state = batch.reset()                                           # [M, 4, N]   environment.py:225-229
    state, reward, done, winner = batch.step(actions)               # actions [M, 2, N]   environment.py:119-223

— and needs no change to bodies.py, geometry.py, the agents or config.py.  The struct mirrors (`cavgym_b200._abi`) and
the walk from `env.bodies / env.constants / env.env_config` to the flat `CavScenario` tables
(`cavgym_b200.scenario.compile_scenario`, which reads the reference's classes by name) ship with the library's Python
package and import nothing but ctypes; tests/test_config.py::test_reference_objects_compile_to_the_same_tables and
tests/test_integration_binding.py run this file against the unmodified reference's own objects.
"""
import ctypes as C

import numpy as np

from cavgym_b200 import _abi, _native
from cavgym_b200.scenario import AgentSpec, compile_scenario


class BatchedCAVEnv:
    """N copies of one reference CAVEnv stepped by the GPU.  Joint actions and results are host numpy arrays in the
    engine's layout (environment index last); `agents` optionally names on-device agents per body ('noop', 'random',
    'random-constrained', 'proximity' — AgentSpec), default: every body takes its action from `step(actions)`."""

    def __init__(self, env, num_envs, seed=0, dtype=np.float64, device=0, agents=None):
        self.lib = _native.load()
        self.tables = compile_scenario(env.bodies, env.constants, env.env_config, agents or [AgentSpec("external") for _ in env.bodies],
                                       time_resolution=env.time_resolution)
        self.n, self.m, self.dtype = int(num_envs), len(env.bodies), np.dtype(dtype)
        self.handle = C.c_void_p()
        code = _abi.CAV_F64 if self.dtype == np.float64 else _abi.CAV_F32
        _native.check(self.lib.cavgym_create(self.tables.pointer(), self.n, code, int(device), int(seed or 0), C.byref(self.handle)))
        # page-locked, device-mapped arrays (cavgym_host_alloc): cavgym_step_host then runs as ONE launch that reads the
        # actions from and writes the results to these arrays over PCIe; pageable arrays work too, through staged copies
        self._pinned = []
        self.actions = self._host_array((self.m, 2, self.n), self.dtype)
        self.state = self._host_array((self.m, 4, self.n), self.dtype)
        self.reward = self._host_array((self.m, self.n), self.dtype)
        self.done = self._host_array((self.n,), np.uint8)
        self.winner = self._host_array((self.n,), np.int32)
        self.tangent = self._host_array((self.n,), np.uint8)
        self.winner[:] = -1

    def _host_array(self, shape, dtype):
        ptr, count = C.c_void_p(), int(np.prod(shape))
        _native.check(self.lib.cavgym_host_alloc(count * np.dtype(dtype).itemsize, 0, C.byref(ptr)))
        self._pinned.append(ptr)
        array = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(np.dtype(dtype)))), shape=(count,)).reshape(shape)
        array[...] = 0
        return array

    @staticmethod
    def _p(array):
        return C.c_void_p(array.ctypes.data)

    def reset(self, mask=None, init_state=None):
        """CAVEnv.reset for the envs selected by `mask` (u8 [N], all when None); SpawnPedestrians are re-drawn on the device
        (bodies.py:299-312) unless `init_state` [M, 4, N] is given.  Returns the observation [M, 4, N]."""
        mask = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        init_state = None if init_state is None else np.ascontiguousarray(init_state, self.dtype)
        _native.check(self.lib.cavgym_reset_host(self.handle, None if mask is None else self._p(mask),
                                                 None if init_state is None else self._p(init_state), self._p(self.state)))
        return self.state

    def step(self, actions):
        """CAVEnv.step over all envs: one fused launch, copies in and out included (cavgym_step_host).  Envs whose joint
        action is invalid (environment.py:120) are left untouched and counted by cavgym_error_count."""
        assert np.shape(actions) == (self.m, 2, self.n), "actions must be [bodies, 2, envs]"
        if actions is not self.actions:
            self.actions[...] = actions          # (write into batch.actions directly to skip this copy)
        _native.check(self.lib.cavgym_step_host(self.handle, self._p(self.actions), self._p(self.state), self._p(self.reward), self._p(self.done),
                                                self._p(self.winner), self._p(self.tangent)))
        return self.state, self.reward, self.done.astype(bool), self.winner

    def stats(self):
        out = (C.c_int64 * _abi.CAV_N_STATS)()
        _native.check(self.lib.cavgym_stats(self.handle, out))
        return dict(zip(_abi.STAT_NAMES, (int(v) for v in out)))

    def close(self):
        if self.handle:
            self.lib.cavgym_destroy(self.handle)
            self.handle = C.c_void_p()
            self.actions = self.state = self.reward = self.done = self.winner = self.tangent = None
            for ptr in self._pinned:
                self.lib.cavgym_host_free(ptr)
            self._pinned = []
