"""BatchedCAVEnv — tensor API over N parallel CAV-Gym environments on one B200.

Host side of the drop-in boundary: the same constructor vocabulary as the reference's
`CAVEnv(bodies, constants, env_config, np_random)` (library/environment.py:59) plus
`num_envs / device / dtype / agents`; `reset()` / `step(actions)` exchange torch tensors in the
engine's SoA layout ([M, 4, N] state, [M, 2, N] actions, [M, N] reward, [N] done / winner).
PyTorch is only plumbing here (device memory, streams); every transition runs in the
hand-written CUDA kernels behind the C-ABI of include/cavgym.h, called through ctypes.
"""
import ctypes as C

import numpy as np
import torch

from . import _abi, _native
from .scenario import AgentSpec, compile_scenario

_DTYPES = {"float64": (_abi.CAV_F64, torch.float64), "float32": (_abi.CAV_F32, torch.float32),
           torch.float64: (_abi.CAV_F64, torch.float64), torch.float32: (_abi.CAV_F32, torch.float32)}


class _DeviceArray:
    """Exposes an engine-owned device buffer through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}
        self._owner = owner


def _ptr(tensor):
    return None if tensor is None else C.c_void_p(tensor.data_ptr())


def _is_float32(array):
    return array is not None and (array.dtype == torch.float32 if isinstance(array, torch.Tensor) else getattr(array, "dtype", None) == np.float32)


def _require_cuda(device):
    if not torch.cuda.is_available():
        raise RuntimeError("cavgym_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if device.type != "cuda":
        raise RuntimeError(f"cavgym_b200 runs on CUDA devices only, not {device}")
    return torch.device("cuda", device.index if device.index is not None else torch.cuda.current_device())


class BatchedCAVEnv:
    def __init__(self, bodies, constants, env_config, num_envs=1, agents=None, device=None, dtype="float64", seed=0,
                 env_offset=0, compiled=None):
        self._lib = _native.load()
        self.device = _require_cuda(device)
        self.code, self.dtype = _DTYPES[dtype]
        self.bodies, self.constants, self.env_config = bodies, constants, env_config
        self.compiled = compiled if compiled is not None else compile_scenario(bodies, constants, env_config, agents)
        self.num_envs, self.num_bodies = int(num_envs), self.compiled.n_bodies
        self.frequency = 60
        self.time_resolution = 1.0 / self.frequency
        self._handle = C.c_void_p()
        _native.check(self._lib.cavgym_create(self.compiled.pointer(), self.num_envs, self.code, self.device.index,
                                              int(seed), C.byref(self._handle)))
        if env_offset:
            _native.check(self._lib.cavgym_set_shard(self._handle, int(env_offset)))
        n, m = self.num_envs, self.num_bodies
        real = "<f8" if self.code == _abi.CAV_F64 else "<f4"
        self.state = self._view(self._lib.cavgym_state_ptr, (m, 4, n), real)
        self.actions_taken = self._view(self._lib.cavgym_action_ptr, (m, 2, n), real)
        self.agent_state = self._view(self._lib.cavgym_agent_state_ptr, (m, _abi.CAV_AGENT_WORDS, n), real)
        self.episode_liveness = self._view(self._lib.cavgym_liveness_ptr, (m, n), "<i4")
        self.timestep = self._view(self._lib.cavgym_timestep_ptr, (n,), "<i4")
        self.done_latch = self._view(self._lib.cavgym_done_ptr, (n,), "|u1")
        self.winner_latch = self._view(self._lib.cavgym_winner_ptr, (n,), "<i4")
        self.error = self._view(self._lib.cavgym_error_ptr, (n,), "|u1")
        self.reward = torch.zeros((m, n), dtype=self.dtype, device=self.device)
        self.done = torch.zeros(n, dtype=torch.uint8, device=self.device)
        self.winner = torch.full((n,), -1, dtype=torch.int32, device=self.device)
        self.tangent = torch.zeros(n, dtype=torch.uint8, device=self.device)
        self._keep = {}
        self._host_seen = {}

    # ---- plumbing ----------------------------------------------------------------------
    def _view(self, getter, shape, typestr):
        return torch.as_tensor(_DeviceArray(getter(self._handle), shape, typestr, self), device=self.device)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _as_real(self, value, shape, name):
        if value is None:
            return None
        tensor = torch.as_tensor(value, dtype=self.dtype, device=self.device).contiguous()
        if tuple(tensor.shape) != tuple(shape):
            raise ValueError(f"{name} must have shape {tuple(shape)}, got {tuple(tensor.shape)}")
        return tensor

    def close(self):
        if getattr(self, "_handle", None):
            self._lib.cavgym_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- reference protocol, batched ------------------------------------------------------
    def reset(self, mask=None, init_state=None):
        """CAVEnv.reset for the envs selected by `mask` (all when None).  `init_state` [M,4,N] replays
        given initial states instead of drawing SpawnPedestrians on the device."""
        n, m = self.num_envs, self.num_bodies
        mask_t = None if mask is None else torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
        init_t = self._as_real(init_state, (m, 4, n), "init_state")
        _native.check(self._lib.cavgym_reset(self._handle, _ptr(mask_t), _ptr(init_t), self._stream()))
        return self.state

    def step(self, actions=None, copy_state=None):
        """CAVEnv.step over all envs.  Returns (state, reward, done, winner, tangent) tensors; `state` is the
        engine's own buffer (zero copy) unless `copy_state` is a [M,4,N] tensor to receive a snapshot."""
        n, m = self.num_envs, self.num_bodies
        actions_t = self._as_real(actions, (m, 2, n), "actions")
        if copy_state is not None:   # written by the kernel through a raw pointer: it must be exactly the engine's layout
            if not (isinstance(copy_state, torch.Tensor) and copy_state.device == self.device and copy_state.dtype == self.dtype
                    and tuple(copy_state.shape) == (m, 4, n) and copy_state.is_contiguous()):
                raise ValueError(f"copy_state must be a contiguous {self.dtype} tensor of shape {(m, 4, n)} on {self.device}")
        _native.check(self._lib.cavgym_step(self._handle, _ptr(actions_t), _ptr(copy_state), _ptr(self.reward), _ptr(self.done),
                                            _ptr(self.winner), _ptr(self.tangent), self._stream()))
        return (self.state if copy_state is None else copy_state), self.reward, self.done, self.winner, self.tangent

    def rollout(self, n_steps, auto_reset=True):
        """Simulation.run's timestep loop with the on-device agents: n_steps fused transitions per launch."""
        _native.check(self._lib.cavgym_rollout(self._handle, int(n_steps), int(bool(auto_reset)), self._stream()))

    def replay(self, actions, record=("state", "reward", "done", "winner", "tangent")):
        """Replay joint actions [T,M,2,N] in one launch; returns the recorded trajectories as a dict."""
        n, m = self.num_envs, self.num_bodies
        actions_t = torch.as_tensor(actions, dtype=self.dtype, device=self.device).contiguous()
        t = actions_t.shape[0]
        if tuple(actions_t.shape) != (t, m, 2, n):
            raise ValueError(f"actions must have shape (T, {m}, 2, {n})")
        out = {
            "state": torch.empty((t, m, 4, n), dtype=self.dtype, device=self.device) if "state" in record else None,
            "reward": torch.empty((t, m, n), dtype=self.dtype, device=self.device) if "reward" in record else None,
            "done": torch.empty((t, n), dtype=torch.uint8, device=self.device) if "done" in record else None,
            "winner": torch.empty((t, n), dtype=torch.int32, device=self.device) if "winner" in record else None,
            "tangent": torch.empty((t, n), dtype=torch.uint8, device=self.device) if "tangent" in record else None,
        }
        _native.check(self._lib.cavgym_replay(self._handle, t, _ptr(actions_t), _ptr(out["state"]), _ptr(out["reward"]),
                                              _ptr(out["done"]), _ptr(out["winner"]), _ptr(out["tangent"]), self._stream()))
        return out

    _NUMPY = {torch.float64: np.dtype("f8"), torch.float32: np.dtype("f4"), torch.uint8: np.dtype("u1"), torch.int32: np.dtype("i4")}

    def _host_buffer(self, array, shape, dtype, name):
        """Raw host pointer of a numpy array / CPU tensor after checking it is what the kernel will read or write."""
        if array is None:
            return None
        seen = self._host_seen.get(id(array))      # validated before: the cache holds the object, so its id cannot be reused
        if seen is not None and seen[0] is array:
            return seen[1]
        pointer = None
        if isinstance(array, torch.Tensor):
            if array.dtype == dtype and array.shape == shape and array.device.type == "cpu" and array.is_contiguous():
                pointer = array.data_ptr()
        elif isinstance(array, np.ndarray) and array.dtype == self._NUMPY[dtype] and array.shape == shape and array.flags.c_contiguous:
            pointer = array.ctypes.data
        if pointer is not None:
            if len(self._host_seen) > 4096:
                self._host_seen.clear()
            self._host_seen[id(array)] = (array, pointer)
            return pointer
        raise ValueError(f"{name} must be a contiguous host array of shape {tuple(shape)} and dtype {dtype}")

    def replay_host(self, actions, state=None, reward=None, done=None, winner=None, tangent=None, wait=True):
        """cavgym_replay with PINNED HOST tensors: T fused steps in one launch whose kernel reads the joint actions [T,M,2,N]
        from and writes the trajectories ([T,M,4,N] state, [T,M,N] reward, [T,N] done / winner / tangent; each optional) to host
        memory over PCIe — reads of later steps overlap writes of earlier ones, which one cavgym_step_host call per step cannot
        do.  Returns when the results are in the host tensors; with wait=False it returns once the launch is queued on the
        current stream (record an event and wait for that before touching the tensors): consecutive calls then overlap on the
        device tile by tile, the reads of one covering the draining writes of the one before."""
        n, m = self.num_envs, self.num_bodies
        t = int(actions.shape[0])
        wanted = (("actions", actions, (t, m, 2, n), self.dtype), ("state", state, (t, m, 4, n), self.dtype),
                  ("reward", reward, (t, m, n), self.dtype), ("done", done, (t, n), torch.uint8), ("winner", winner, (t, n), torch.int32),
                  ("tangent", tangent, (t, n), torch.uint8))
        pointers = []
        for name, tensor, shape, dtype in wanted:
            if tensor is None:
                pointers.append(None)
                continue
            if not (isinstance(tensor, torch.Tensor) and tensor.device.type == "cpu" and tensor.is_pinned() and tensor.dtype == dtype
                    and tuple(tensor.shape) == shape and tensor.is_contiguous()):
                raise ValueError(f"{name} must be a pinned, contiguous CPU tensor of shape {shape} and dtype {dtype}")
            pointers.append(C.c_void_p(tensor.data_ptr()))
        stream = torch.cuda.current_stream(self.device)
        _native.check(self._lib.cavgym_replay(self._handle, t, *pointers, C.c_void_p(stream.cuda_stream)))
        if wait:
            stream.synchronize()

    def step_host(self, actions, state_out=None, reward_out=None, done_out=None, winner_out=None, tangent_out=None):
        """cavgym_step_host: host buffers in and out (numpy arrays or CPU tensors in the engine's layout).  Pinned buffers
        (tensor.pin_memory()) are read and written by the step kernel itself over PCIe; pageable ones are staged.
        A float64 engine also takes float32 actions / state_out / reward_out (all three the same type): the float32 WIRE
        format of cavgym_step_host_f32 — half the bytes over the link, the engine still steps in double."""
        n, m = self.num_envs, self.num_bodies
        check = self._host_buffer
        real = self.dtype
        if self.dtype == torch.float64 and _is_float32(actions if actions is not None else (state_out if state_out is not None else reward_out)):
            real = torch.float32
        call = self._lib.cavgym_step_host if real == self.dtype else self._lib.cavgym_step_host_f32
        rc = call(self._handle, check(actions, (m, 2, n), real, "actions"), check(state_out, (m, 4, n), real, "state_out"),
                  check(reward_out, (m, n), real, "reward_out"), check(done_out, (n,), torch.uint8, "done_out"),
                  check(winner_out, (n,), torch.int32, "winner_out"), check(tangent_out, (n,), torch.uint8, "tangent_out"))
        if rc:
            _native.check(rc)

    def info(self, polygons=True, road_angles=True):
        """CAVEnv.info() for every env: {'body_polygons': [M, 8, N] (x of the four corners, then y), 'road_angles': [M, N]
        with NaN where the reference returns None}.  Computed on demand by one small kernel."""
        n, m = self.num_envs, self.num_bodies
        out = {"body_polygons": torch.empty((m, 8, n), dtype=self.dtype, device=self.device) if polygons else None,
               "road_angles": torch.empty((m, n), dtype=self.dtype, device=self.device) if road_angles else None}
        _native.check(self._lib.cavgym_info(self._handle, _ptr(out["body_polygons"]), _ptr(out["road_angles"]), self._stream()))
        return out

    # ---- accounting and knobs ------------------------------------------------------------
    def stats(self):
        out = (C.c_int64 * _abi.CAV_N_STATS)()
        _native.check(self._lib.cavgym_stats(self._handle, out))
        return dict(zip(_abi.STAT_NAMES, [int(v) for v in out]))

    def set_episode_log(self, capacity):
        """Keep a device ring of `capacity` per-episode rows (0: off); see drain_episodes."""
        _native.check(self._lib.cavgym_set_episode_log(self._handle, int(capacity)))
        self._episode_capacity = int(capacity)

    def drain_episodes(self):
        """(rows, dropped): the episodes scored since the last drain as a numpy structured array with the fields of
        CavEpisodeRow (env, episode, timesteps, winner, liveness_sum), in the order they finished, and how many rows a
        full ring lost."""
        capacity = getattr(self, "_episode_capacity", 0)
        rows = np.zeros(capacity, dtype=np.dtype([("env", "<i8"), ("episode", "<i4"), ("timesteps", "<i4"), ("winner", "<i4"),
                                                  ("liveness_sum", "<i4")]))
        n_rows, dropped = C.c_int64(), C.c_int64()
        _native.check(self._lib.cavgym_drain_episodes(self._handle, C.c_void_p(rows.ctypes.data), capacity, C.byref(n_rows), C.byref(dropped)))
        return rows[:n_rows.value], int(dropped.value)

    def launch_count(self):
        out = C.c_int64()
        _native.check(self._lib.cavgym_launch_count(self._handle, C.byref(out)))
        return int(out.value)

    def set_global_timestep(self, t):
        _native.check(self._lib.cavgym_set_global_timestep(self._handle, int(t)))

    def set_step_path(self, use_tma=True):
        """False forces the plain thread-per-env step kernel (the TMA-staged kernel is the default where it applies)."""
        _native.check(self._lib.cavgym_set_step_path(self._handle, int(bool(use_tma))))

    def set_rollout_path(self, team=True):
        """True runs cavgym_rollout of a heterogeneous scenario of three or more bodies on the team-of-warps kernel
        (kernels_team.cuh: one warp per body; bitwise the thread-per-env kernel, which stays the default)."""
        _native.check(self._lib.cavgym_set_rollout_path(self._handle, int(bool(team))))

    def set_dense_path(self, force=True):
        """True runs a small scenario through the warp-per-env kernels that scenarios with > CAV_SMALL_M bodies always use."""
        _native.check(self._lib.cavgym_set_dense_path(self._handle, int(bool(force))))

    def set_host_path(self, zero_copy=True):
        """False makes step_host stage pinned buffers through device copies instead of the zero-copy launch."""
        _native.check(self._lib.cavgym_set_host_path(self._handle, int(zero_copy)))

    def set_tangent_tolerance(self, tau):
        _native.check(self._lib.cavgym_set_tangent_tolerance(self._handle, float(tau)))

    def set_action_logging(self, enabled=True):
        _native.check(self._lib.cavgym_set_action_logging(self._handle, int(bool(enabled))))

    def set_uniform_override(self, uniforms):
        self._keep["uniforms"] = None if uniforms is None else torch.as_tensor(
            uniforms, dtype=torch.float64, device=self.device).contiguous()
        _native.check(self._lib.cavgym_set_uniform_override(self._handle, _ptr(self._keep["uniforms"])))

    def set_spawn_override(self, draws):
        self._keep["spawn"] = None if draws is None else torch.as_tensor(draws, dtype=torch.float64, device=self.device).contiguous()
        _native.check(self._lib.cavgym_set_spawn_override(self._handle, _ptr(self._keep["spawn"])))

    # ---- single-environment helpers for the compat view (library/environment.py) ----------------
    def reset_to(self, rows):
        init = torch.tensor(rows, dtype=self.dtype).reshape(self.num_bodies, 4, 1).expand(-1, -1, self.num_envs)
        self.reset(init_state=init.contiguous())

    def step_single(self, joint_action):
        rows = [[float(a[0]), float(a[1])] if isinstance(a, (list, tuple, np.ndarray)) else [float(int(a)), 0.0] for a in joint_action]
        actions = torch.tensor(rows, dtype=self.dtype).reshape(self.num_bodies, 2, 1)
        state, reward, done, winner, _ = self.step(actions)
        packed = torch.cat([state[:, :, 0].reshape(-1).double(), reward[:, 0].double(), done[:1].double(), winner[:1].double(),
                            self.episode_liveness[:, 0].double()]).cpu().tolist()
        m = self.num_bodies
        state_rows = [packed[4 * b:4 * b + 4] for b in range(m)]
        rewards = packed[4 * m:5 * m]
        done_flag, winner_index = bool(packed[5 * m]), int(packed[5 * m + 1])
        liveness = [int(v) for v in packed[5 * m + 2:6 * m + 2]]
        taken = [[r[0], 0.0 if abs(r[1]) < 0.0000000000001 else r[1]] for r in rows]
        return state_rows, rewards, done_flag, winner_index, liveness, taken


class HostBuffer:
    """A page-locked, device-mapped host array from cavgym_host_alloc, as a numpy array (`.array`).  write_combined=True is for
    buffers the host only writes and the GPU reads — the joint actions of step_host."""

    def __init__(self, shape, dtype, write_combined=False):
        self._lib = _native.load()
        self.array = None
        count = int(np.prod(shape))
        self._ptr = C.c_void_p()
        _native.check(self._lib.cavgym_host_alloc(count * np.dtype(dtype).itemsize, int(bool(write_combined)), C.byref(self._ptr)))
        ctype = np.ctypeslib.as_ctypes_type(np.dtype(dtype))
        self.array = np.ctypeslib.as_array(C.cast(self._ptr, C.POINTER(ctype)), shape=(count,)).reshape(shape)

    def close(self):
        if getattr(self, "_ptr", None) and self._ptr.value:
            self.array = None
            self._lib.cavgym_host_free(self._ptr)
            self._ptr = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def bodies_step(constants, states, actions, time_resolution, dtype="float64", device=None):
    """DynamicBody.step (reference library/bodies.py:214-275) for a list of independent bodies of one type,
    run by the stand-alone CUDA kinematics kernel.  states [n][4], actions [n][2] -> new states [n][4]."""
    lib = _native.load()
    device = _require_cuda(device)
    code, tdtype = _DTYPES[dtype]
    state_t = torch.tensor(states, dtype=tdtype, device=device).t().contiguous()
    action_t = torch.tensor(actions, dtype=tdtype, device=device).t().contiguous()
    k = _abi.CavBodyType(*[float(v) for v in (constants.length, constants.width, constants.wheelbase, constants.min_velocity,
                                               constants.max_velocity, constants.min_throttle, constants.max_throttle,
                                               constants.min_steering_angle, constants.max_steering_angle)])
    with torch.cuda.device(device):
        _native.check(lib.cavgym_bodies_step(C.byref(k), _ptr(state_t), _ptr(action_t), state_t.shape[1], float(time_resolution),
                                             code, C.c_void_p(torch.cuda.current_stream(device).cuda_stream)))
    return state_t.t().cpu().tolist()


def zones_probe(constants, states, steering, dtype="float64", device=None):
    """DynamicBody.stopping_zones (reference library/bodies.py:122-135) for n independent bodies of one type, as the step
    kernels represent the zones.  states [n][4], steering [n] -> (zones [n][2][4][2] = braking / reaction corner lists in the
    reference's order, have [n] bool; rows without zones are NaN)."""
    lib = _native.load()
    device = _require_cuda(device)
    code, tdtype = _DTYPES[dtype]
    state_t = torch.tensor(np.asarray(states, dtype=np.float64), dtype=tdtype, device=device).t().contiguous()
    steer_t = torch.tensor(np.asarray(steering, dtype=np.float64), dtype=tdtype, device=device).contiguous()
    n = state_t.shape[1]
    zones = torch.empty((16, n), dtype=tdtype, device=device)
    have = torch.empty(n, dtype=torch.uint8, device=device)
    k = _abi.CavBodyType(*[float(v) for v in (constants.length, constants.width, constants.wheelbase, constants.min_velocity,
                                               constants.max_velocity, constants.min_throttle, constants.max_throttle,
                                               constants.min_steering_angle, constants.max_steering_angle)])
    with torch.cuda.device(device):
        _native.check(lib.cavgym_zones_probe(C.byref(k), _ptr(state_t), _ptr(steer_t), _ptr(zones), _ptr(have), n, code,
                                             C.c_void_p(torch.cuda.current_stream(device).cuda_stream)))
    z = zones.double().cpu().numpy().reshape(2, 2, 4, n)           # [zone][x|y][corner][n]
    return np.transpose(z, (3, 0, 2, 1)), have.cpu().numpy().astype(bool)


def geometry_probe(quads_a, quads_b, dtype="float64", device=None):
    """Shape.intersects / contains / percentage_intersects (reference library/geometry.py:74-87) on n quad pairs,
    run by the CUDA geometry code.  quads [n][4][2] -> array [n][4]: intersects, b.contains(a), share, tangent."""
    lib = _native.load()
    device = _require_cuda(device)
    code, tdtype = _DTYPES[dtype]

    def pack(quads):
        arr = np.asarray(quads, dtype=np.float64)  # [n, 4, 2]
        return torch.tensor(np.concatenate([arr[:, :, 0].T, arr[:, :, 1].T], axis=0), dtype=tdtype, device=device).contiguous()

    a, b = pack(quads_a), pack(quads_b)
    out = torch.empty((4, a.shape[1]), dtype=tdtype, device=device)
    with torch.cuda.device(device):
        _native.check(lib.cavgym_geometry_probe(_ptr(a), _ptr(b), _ptr(out), a.shape[1], code,
                                                C.c_void_p(torch.cuda.current_stream(device).cuda_stream)))
    return out.t().double().cpu().numpy()
