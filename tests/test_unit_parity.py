"""Per-function parity (SURVEY §4 "per-function unit parity") against what the UNMODIFIED reference computed:

  SpawnPedestrian.spawn      bodies.py:302-312, geometry.py:223-229, 369-377   fixtures' `spawn_draws` -> `init_state`
  Shape.intersects / contains / percentage_intersects   geometry.py:74-87     geometry_kat.npz['pairs']
  DynamicBody.bounding_box   bodies.py:116-117                                 geometry_kat.npz['pairs'] corner lists
  DynamicBody.stopping_zones bodies.py:122-135, geometry.py:176-191            geometry_kat.npz['zones']
  DynamicBody.step           bodies.py:214-275                                 oracle restatement (itself bit-equal on the traces)

CPU tests pin the oracle's single-shot helpers to the fixtures; the `gpu` tests run the engine's stand-alone hooks
(cavgym_set_spawn_override + cavgym_reset, cavgym_geometry_probe, cavgym_zones_probe, cavgym_bodies_step) through the C-ABI.
"""
import ctypes as C
import math
import os

import numpy as np
import pytest

from helpers import GOLDEN_CASES, GOLDEN_DIR, compile_from_meta, load_golden, soa, state_err
from cavgym_b200 import _abi
from oracle import oracle as orc

KAT = np.load(os.path.join(GOLDEN_DIR, "geometry_kat.npz"))
SPAWN_CASES = [name for name in GOLDEN_CASES if name.startswith("pedestrians")]
REL = {"float64": 1e-9, "float32": 1e-4}


def spawn_override(meta, episodes):
    """[M, 5, N] spawn draws of the reference, one env per recorded episode (bodies without a spawner: unused zeros)."""
    m, n = meta["n_bodies"], len(episodes)
    draws = np.zeros((m, 5, n))
    for e, ep in enumerate(episodes):
        flat = np.asarray(ep["spawn_draws"], dtype=np.float64).reshape(-1, 5)
        assert flat.shape[0] == m - 1           # the pedestrians scenarios: one SpawnPedestrian per non-ego body
        draws[1:, :, e] = flat
    return draws


# ------------------------------------------------------------------------------------------------ CPU: the oracle
@pytest.mark.parametrize("name", SPAWN_CASES)
def test_oracle_spawn_matches_reference_draws(name):
    """Fed the uniforms the reference's shared RandomState produced, the oracle's sampler lands on the reference's
    post-reset state bit for bit (area-weighted box, area-weighted triangle, reflected point, orientation)."""
    meta, episodes = load_golden(name)
    sim = orc.Oracle(compile_from_meta(meta), len(episodes))
    sim.set_spawn_override(spawn_override(meta, episodes))
    sim.reset()
    want = soa(np.stack([ep["init_state"] for ep in episodes]))
    assert np.array_equal(sim.state, want)


def test_oracle_shapely_bridge_matches_reference_kat():
    lib = orc.lib()
    rows = KAT["pairs"]
    worst = 0.0
    for row in rows:
        la, wa, xa, ya, ta, lb, wb, xb, yb, tb, hit, inside, share = row[:13]
        a, b = _abi.CavQuad(), _abi.CavQuad()
        lib.cav_oracle_make_box(la, wa, ta, xa, ya, C.byref(a))
        lib.cav_oracle_make_box(lb, wb, tb, xb, yb, C.byref(b))
        assert [c for p in orc.quad_points(a) for c in p] == row[13:21].tolist()    # bounding_box corners, bit-equal
        assert [c for p in orc.quad_points(b) for c in p] == row[21:29].tolist()
        assert lib.cav_oracle_intersects(C.byref(a), C.byref(b)) == int(hit)
        assert lib.cav_oracle_contains(C.byref(b), C.byref(a)) == int(inside)
        worst = max(worst, abs(lib.cav_oracle_percentage_intersects(C.byref(a), C.byref(b)) - share))
    assert worst < 1e-12
    assert rows[:, 10].any() and not rows[:, 10].all()      # the vectors exercise both answers


def test_oracle_stopping_zones_match_reference_kat():
    lib = orc.lib()
    some = none = 0
    for row in KAT["zones"]:
        x, y, v, th, length, width, min_throttle = row[:7]
        k = _abi.CavBodyType(length, width, 0.0, 0.0, 0.0, min_throttle, 0.0, 0.0, 0.0)
        st = (C.c_double * 4)(x, y, v, th)
        braking, reaction = _abi.CavQuad(), _abi.CavQuad()
        have = lib.cav_oracle_stopping_zones(C.byref(k), st, 0.0, C.byref(braking), C.byref(reaction))
        if np.isnan(row[7]):
            assert have == 0
            none += 1
            continue
        assert have == 1
        got = [c for p in orc.quad_points(braking) for c in p] + [c for p in orc.quad_points(reaction) for c in p]
        assert got == row[7:23].tolist()                                             # bit-equal corner lists
        assert lib.cav_oracle_stopping_zones(C.byref(k), st, 0.1, C.byref(braking), C.byref(reaction)) == 0   # steering: None
        some += 1
    assert some and none


# ------------------------------------------------------------------------------------------------ GPU: the engine's hooks
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("name", SPAWN_CASES)
def test_engine_spawn_matches_reference_draws(name, dtype):
    """cavgym_set_spawn_override + cavgym_reset: the device sampler on the reference's draws reproduces the reference's
    post-reset state (fp64: exactly — the sampler works in double; fp32: to float32 rounding of the spawn boxes)."""
    from cavgym_b200 import BatchedCAVEnv
    meta, episodes = load_golden(name)
    env = BatchedCAVEnv(None, None, None, num_envs=len(episodes), dtype=dtype, compiled=compile_from_meta(meta))
    env.set_spawn_override(spawn_override(meta, episodes))
    env.reset()
    got = env.state.double().cpu().numpy()
    want = soa(np.stack([ep["init_state"] for ep in episodes]))
    if dtype == "float64":
        assert np.array_equal(got, want)
    else:   # the spawn boxes themselves are float32 there
        assert state_err(np.moveaxis(got, 1, -1), np.moveaxis(want, 1, -1)) < 1e-6
    env.set_spawn_override(None)
    env.close()


def kat_quads(rows, first):
    return rows[:, first:first + 8].reshape(-1, 4, 2)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_engine_geometry_probe_matches_reference_kat(dtype):
    """cavgym_geometry_probe on the corner lists the reference built: predicates equal wherever the engine does not raise
    its near-tangent flag (exact touching is flagged by construction), share within 1e-9 / 1e-4."""
    from cavgym_b200.engine import geometry_probe
    rows = KAT["pairs"]
    out = geometry_probe(kat_quads(rows, 13), kat_quads(rows, 21), dtype=dtype)
    clear = out[:, 3] == 0
    assert clear.sum() > 0.6 * len(rows)
    assert np.array_equal(out[clear, 0], rows[clear, 10])
    assert np.array_equal(out[clear, 1], rows[clear, 11])
    assert np.max(np.abs(out[clear, 2] - rows[clear, 12])) < REL[dtype]
    # flagged rows: the share is continuous across a tangency, so it still has to agree
    assert np.max(np.abs(out[~clear, 2] - rows[~clear, 12])) < (1e-6 if dtype == "float64" else 2e-2)
    touching = np.arange(len(rows)) % 4 == 1        # the generator's "exactly touching edges" rows
    assert not clear[touching].all()                # some exact tangencies exist and are flagged


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_engine_stopping_zones_match_reference_kat(dtype):
    """cavgym_zones_probe: the rectangles the step kernels test pedestrians against are the reference's zones."""
    from types import SimpleNamespace
    from cavgym_b200.engine import zones_probe
    rows = KAT["zones"]
    for length in np.unique(rows[:, 4]):
        sel = rows[rows[:, 4] == length]
        k = SimpleNamespace(length=float(length), width=float(sel[0, 5]), wheelbase=1.0, min_velocity=0.0, max_velocity=0.0,
                            min_throttle=float(sel[0, 6]), max_throttle=0.0, min_steering_angle=-1.0, max_steering_angle=1.0)
        zones, have = zones_probe(k, sel[:, :4], np.zeros(len(sel)), dtype=dtype)
        want_have = ~np.isnan(sel[:, 7])
        assert np.array_equal(have, want_have)
        want = sel[want_have, 7:23].reshape(-1, 2, 4, 2)
        got = zones[want_have]
        scale = np.maximum(1.0, np.abs(want))
        assert np.max(np.abs(got - want) / scale) < REL[dtype]
        _, steering = zones_probe(k, sel[:, :4], np.full(len(sel), 0.05), dtype=dtype)
        assert not steering.any()                   # zones vanish while the body steers (bodies.py:130-135)
        _, snapped = zones_probe(k, sel[:, :4], np.full(len(sel), 5e-14), dtype=dtype)
        assert np.array_equal(snapped, want_have)   # |steer| < 1e-13 is no steering (bodies.py:217-218)


def step_cases(rs, k, n):
    """Random (state, action) rows for one body type with the special cases of DynamicBody.step."""
    state = np.stack([rs.uniform(-100, 1700, n), rs.uniform(-120, 120, n), rs.uniform(k.min_velocity, k.max_velocity, n),
                      rs.uniform(-math.pi, math.pi, n)], axis=1)
    action = np.stack([rs.uniform(k.min_throttle, k.max_throttle, n), rs.uniform(k.min_steering_angle, k.max_steering_angle, n)], axis=1)
    kind = rs.randint(0, 10, n)
    action[kind == 0, 1] = 0.0                                           # straight
    action[kind == 1, 1] = rs.uniform(-1e-13, 1e-13, (kind == 1).sum())  # snapped to straight
    action[kind == 2, 1] = k.max_steering_angle                          # full lock
    action[kind == 3, 1] = k.min_steering_angle
    action[kind == 4, 1] = rs.uniform(-1e-6, 1e-6, (kind == 4).sum())    # tiny steering: huge turn radius
    state[kind == 5, 2] = 0.0                                            # standing still
    state[kind == 6, 3] = rs.choice([0.0, math.pi, -math.pi, math.pi / 2, -math.pi / 2], (kind == 6).sum())
    state[kind == 7, 2] = k.max_velocity                                 # clamped at either end
    action[kind == 7, 0] = k.max_throttle
    state[kind == 8, 2] = k.min_velocity
    action[kind == 8, 0] = k.min_throttle
    return state, action


def oracle_step(k, state, action, dt):
    lib = orc.lib()
    kt = _abi.CavBodyType(*[float(v) for v in (k.length, k.width, k.wheelbase, k.min_velocity, k.max_velocity, k.min_throttle,
                                               k.max_throttle, k.min_steering_angle, k.max_steering_angle)])
    out = np.empty_like(state)
    st = (C.c_double * 4)()
    for i in range(len(state)):
        st[:] = state[i]
        lib.cav_oracle_dynamic_body_step(C.byref(kt), st, float(action[i, 0]), float(action[i, 1]), dt)
        out[i] = st[:]
    return out


def turn_exact(k, state, action, dt):
    """DynamicBody.step's turning branch (bodies.py:244-275) on the body-frame offset d = (wb/2)(c, s) + K(s, -c) from the
    centre of rotation: p' = p + d (cos phi - 1) + d_perp sin phi, with cos phi - 1 = -2 sin^2(phi / 2).  Algebraically the
    reference's formula, but nothing of size K is ever subtracted from anything of size K."""
    x, y, v, th = state.T
    c, s = np.cos(th), np.sin(th)
    big_k = k.wheelbase / np.tan(action[:, 1])
    dx, dy = 0.5 * k.wheelbase * c + big_k * s, 0.5 * k.wheelbase * s - big_k * c
    phi = np.sign(action[:, 1]) * (v * dt) / np.hypot(dx, dy)
    cm1, sn = -2.0 * np.sin(0.5 * phi) ** 2, np.sin(phi)
    out = np.empty_like(state)
    out[:, 0] = x + (dx * cm1 - dy * sn)
    out[:, 1] = y + (dx * sn + dy * cm1)
    out[:, 2] = np.clip(v + action[:, 0] * dt, k.min_velocity, k.max_velocity)
    out[:, 3] = np.arctan2(np.sin(th + phi), np.cos(th + phi))
    return out


def test_turn_exact_is_the_reference_turn_where_the_reference_is_accurate():
    """turn_exact against the oracle's literal restatement wherever the reference's cancellation is harmless."""
    from cavgym_b200.examples.constants import car_constants, pedestrian_constants
    rs = np.random.RandomState(3)
    for k in (car_constants, pedestrian_constants):
        state, action = step_cases(rs, k, 20000)
        turning = np.abs(action[:, 1]) > 1e-3
        assert state_err(turn_exact(k, state[turning], action[turning], 1.0 / 60), oracle_step(k, state[turning], action[turning], 1.0 / 60)) < 1e-11


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_engine_bodies_step_matches_oracle(dtype):
    """cavgym_bodies_step (the Body.step plugin hook) on 100,000 random bodies per type against DynamicBody.step as the
    oracle restates it, including |steer| < 1e-13, full lock, v = 0 and headings of 0, +-pi/2, +-pi."""
    from cavgym_b200.engine import bodies_step
    from cavgym_b200.examples.constants import car_constants, pedestrian_constants
    rs = np.random.RandomState(11)
    dt = 1.0 / 60
    for k in (car_constants, pedestrian_constants):
        state, action = step_cases(rs, k, 100000)
        if dtype == "float32":   # the same inputs in both engines: what float32 can hold
            state, action = state.astype(np.float32).astype(np.float64), action.astype(np.float32).astype(np.float64)
            action[np.abs(action[:, 1]) < 1e-13, 1] = 0.0
        got = np.asarray(bodies_step(k, state.tolist(), action.tolist(), dt, dtype=dtype))
        want = oracle_step(k, state, action, dt)
        # bodies.py:244-262 builds the centre of rotation in WORLD coordinates and rotates about it: two numbers of size
        # K = wheelbase / tan(steer) are subtracted, so the reference's own result carries an error of ~eps * K pixels
        # (K = 4e12 px at |steer| = 1e-12).  Rows where that exceeds a tenth of the tolerance are judged against the same
        # turn written without the cancellation (turn_exact below); the reference must then agree within ITS error bound.
        radius = np.abs(k.wheelbase / np.tan(np.where(np.abs(action[:, 1]) < 1e-13, 1.0, action[:, 1])))
        radius[np.abs(action[:, 1]) < 1e-13] = 0.0
        scale = np.maximum(1.0, np.hypot(want[:, 0], want[:, 1]))
        cancels = 16 * np.finfo(np.float64).eps * radius > 0.1 * REL[dtype] * scale
        assert state_err(got[~cancels], want[~cancels]) < REL[dtype]
        if cancels.any():
            exact = turn_exact(k, state[cancels], action[cancels], dt)
            assert state_err(got[cancels], exact) < REL[dtype]
            slack = np.hypot(want[cancels, 0] - exact[:, 0], want[cancels, 1] - exact[:, 1])
            assert np.all(slack <= 16 * np.finfo(np.float64).eps * radius[cancels] + 1e-9 * scale[cancels])
            assert cancels.mean() < 0.12   # only the tiny-steering rows (case 4) can be in this class
        straight = np.abs(action[:, 1]) < 1e-13
        assert np.array_equal(got[straight, 3], want[straight, 3])           # heading untouched when not steering
        assert np.all(np.abs(got[:, 3]) <= math.pi + 1e-6)


@pytest.mark.gpu
def test_compat_dynamic_body_step_runs_the_kernel():
    """library.bodies.DynamicBody.step (the compat plugin surface) is the same kernel: one body, one action."""
    from cavgym_b200.examples.constants import pedestrian_constants as k
    from cavgym_b200.library import bodies
    from cavgym_b200.library.geometry import Point
    rs = np.random.RandomState(5)
    state, action = step_cases(rs, k, 40)
    want = oracle_step(k, state, action, 1.0 / 60)
    for i in range(len(state)):
        body = bodies.Pedestrian(bodies.DynamicBodyState(Point(state[i, 0], state[i, 1]), state[i, 2], state[i, 3]), k)
        body.step(action[i].tolist(), 1.0 / 60)
        got = np.array([body.state.position.x, body.state.position.y, body.state.velocity, body.state.orientation])
        assert state_err(got[None], want[i][None]) < 1e-9
