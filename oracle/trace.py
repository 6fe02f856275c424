"""Trace recorder for the CAV-Gym reference (TEST INFRASTRUCTURE ONLY, build container only).

Drives the reference's own `Simulation.run` (simulation.py:38-118) on an env and
agents built by the reference's own `Config.setup` (config.py:272-415) and
records, per episode: post-reset state, and per timestep the joint action,
post-step state, joint reward, done, winner, episode_liveness, the [0,1) draws
each agent consumed, and each CrossingAgent's internal state after
`process_feedback`.  The arrays are what tests/golden/*.npz hold.
"""
import copy
import math

import numpy as np

from . import refload

NAN = float("nan")


class RecordingRandomState:
    """Wraps numpy's legacy RandomState.  Each method restates numpy's legacy algorithm
    in terms of `random_sample`/`randint`, logs the primitive draws, and asserts
    bit-equality against a twin generator running the real numpy method."""

    def __init__(self, rs):
        self._rs = rs
        self._twin = np.random.RandomState()
        self._twin.set_state(rs.get_state())
        self.log = []

    def random_sample(self, size=None):
        u = self._rs.random_sample(size)
        assert np.array_equal(u, self._twin.random_sample(size))
        self.log.append(("u", np.atleast_1d(u).astype(float).tolist()))
        return u

    def uniform(self, low=0.0, high=1.0, size=None):
        u = self._rs.random_sample(size)
        value = low + (high - low) * u
        assert np.array_equal(value, self._twin.uniform(low, high, size))
        self.log.append(("u", np.atleast_1d(u).astype(float).tolist()))
        return value

    def randint(self, low, high=None, size=None):
        value = self._rs.randint(low, high, size)
        assert np.array_equal(value, self._twin.randint(low, high, size))
        lo, hi = (0, low) if high is None else (low, high)
        self.log.append(("i", [(float(v) - lo + 0.5) / (hi - lo) for v in np.atleast_1d(value)]))
        return value

    def choice(self, a, size=None, replace=True, p=None):
        assert size is None and replace
        expected = self._twin.choice(a, p=p)
        n = len(a)
        if p is None:
            idx = int(self._rs.randint(0, n))
            self.log.append(("i", [(idx + 0.5) / n]))
        else:
            cdf = np.cumsum(np.asarray(p, dtype=float))
            cdf /= cdf[-1]
            u = self._rs.random_sample()
            idx = int(cdf.searchsorted(u, side='right'))
            self.log.append(("u", [float(u)]))
        value = a[idx]
        assert value is expected or value == expected
        return value

    def normal(self, loc=0.0, scale=1.0, size=None):
        assert size is not None and int(np.prod(size)) == 0, "unbounded Box dims are not on the traced path"
        self._twin.normal(loc, scale, size)
        return self._rs.normal(loc, scale, size)

    def exponential(self, scale=1.0, size=None):
        assert size is not None and int(np.prod(size)) == 0, "half-bounded Box dims are not on the traced path"
        self._twin.exponential(scale, size)
        return self._rs.exponential(scale, size)

    def flat(self, start):
        out = []
        for _, values in self.log[start:]:
            out.extend(values)
        return out


def _opt(x):
    return NAN if x is None else float(x)


def _agent_state(agent):
    """CrossingAgent internals (examples/agents/pedestrian.py:14-31); NaN encodes None."""
    if hasattr(agent, "waypoint"):
        wp = agent.waypoint
        return [_opt(agent.initial_distance), _opt(None if wp is None else wp.x), _opt(None if wp is None else wp.y),
                _opt(agent.target_orientation), _opt(agent.prior_orientation)]
    return [NAN] * 5


def _action_row(action):
    if isinstance(action, (list, tuple, np.ndarray)):
        return [float(action[0]), float(action[1])]
    return [float(int(action)), 0.0]  # PelicanCrossing: TrafficLightAction value


def _state_row(body_state):
    row = [float(v) if not hasattr(v, "value") else float(v.value) for v in body_state]
    return row + [0.0] * (4 - len(row))  # PelicanCrossing state = [TrafficLightState]


def record(config_dict, max_draws=3, extra=None):
    """Run the reference on `config_dict`; return (meta, [episode dicts of numpy arrays]).
    `extra(env, agents, simulation)` (optional) returns a flat list of floats sampled after the LAST agent's
    process_feedback of every step (learning weights, election flags ...): stored as `extra` [T, K]."""
    mods = refload.load()
    from gym.utils import seeding

    rngs = []

    def wrap(rs):
        rec = RecordingRandomState(rs)
        rngs.append(rec)
        return rec

    seeding.set_rng_wrapper(wrap)
    try:
        cfg = mods["config"].make_config(copy.deepcopy(config_dict))  # make_config pops keys
        _, env, agents, keyboard_agent = cfg.setup()
    finally:
        seeding.set_rng_wrapper(None)
    rng = env.np_random
    assert isinstance(rng, RecordingRandomState)
    # A RandomAgent ego is built without np_random (config.py:305-310, quirk 9): give it a recorded one.
    for agent in agents:
        own = getattr(agent, "np_random", None)
        if own is not None and not isinstance(own, RecordingRandomState):
            own = np.random.RandomState(10_000 + int(config_dict.get("seed") or 0))  # reproducible fixtures
            agent.np_random = RecordingRandomState(own)

    n_bodies = len(env.bodies)
    episodes = []
    current = {}
    holder = {}

    def begin_episode(state, spawn_draws, t_global):
        current.clear()
        current.update(init_state=[_state_row(s) for s in state], spawn_draws=spawn_draws, t_global_start=t_global,
                       actions=[], state=[], reward=[], done=[], winner=[], liveness=[], draws=[], agent_state=[], extra=[])

    pending_draws = [[NAN] * max_draws for _ in range(n_bodies)]

    def wrap_agent(index, agent):
        choose, feedback = agent.choose_action, agent.process_feedback
        agent_rng = getattr(agent, "np_random", None)

        def choose_action(state, action_space, info=None):
            sources = [agent_rng, rng] if agent_rng is not None and agent_rng is not rng else [rng]
            marks = [(r, len(r.log)) for r in sources]  # the epsilon draw precedes the space sample
            action = choose(state, action_space, info)
            used = []
            for r, start in marks:
                used.extend(r.flat(start))
            assert len(used) <= max_draws, used
            pending_draws[index] = used + [NAN] * (max_draws - len(used))
            return action

        def process_feedback(previous_state, action, state, reward):
            feedback(previous_state, action, state, reward)
            current["agent_state"][-1][index] = _agent_state(agent)
            if extra is not None and index == n_bodies - 1:
                current["extra"].append([float(v) for v in extra(env, agents, holder["simulation"])])

        agent.choose_action = choose_action
        agent.process_feedback = process_feedback

    for index, agent in enumerate(agents):
        wrap_agent(index, agent)

    env_step, env_reset = env.step, env.reset

    def reset():
        if current:
            episodes.append({k: v for k, v in current.items()})
        start = len(rng.log)
        state = env_reset()
        begin_episode(state, rng.flat(start), env.current_timestep)
        return state

    def step(joint_action):
        state, joint_reward, done, info = env_step(joint_action)
        current["actions"].append([_action_row(a) for a in joint_action])
        current["state"].append([_state_row(s) for s in state])
        current["reward"].append([float(r) for r in joint_reward])
        current["done"].append(bool(done))
        current["winner"].append(int(info["winner"]) if "winner" in info else -1)
        current["liveness"].append(list(env.episode_liveness))
        current["draws"].append([list(d) for d in pending_draws])
        current["agent_state"].append([[NAN] * 5 for _ in range(n_bodies)])
        return state, joint_reward, done, info

    env.reset, env.step = reset, step
    holder["simulation"] = mods["simulation"].Simulation(env, agents, config=cfg, keyboard_agent=keyboard_agent)
    holder["simulation"].run()
    episodes.append({k: v for k, v in current.items()})

    bodies_mod = mods["bodies"]
    meta = {
        "config": config_dict,
        "body_classes": [type(b).__name__ for b in env.bodies],
        "agent_classes": [type(a).__name__ for a in agents],
        "n_bodies": n_bodies,
        "is_pedestrian": [isinstance(b, bodies_mod.Pedestrian) for b in env.bodies],
    }
    out = []
    for ep in episodes:
        arrays = {
            "init_state": np.asarray(ep["init_state"], dtype=np.float64),
            "spawn_draws": np.asarray(ep["spawn_draws"], dtype=np.float64),
            "t_global_start": np.asarray(ep["t_global_start"], dtype=np.int64),
            "actions": np.asarray(ep["actions"], dtype=np.float64),
            "state": np.asarray(ep["state"], dtype=np.float64),
            "reward": np.asarray(ep["reward"], dtype=np.float64),
            "done": np.asarray(ep["done"], dtype=np.uint8),
            "winner": np.asarray(ep["winner"], dtype=np.int32),
            "liveness": np.asarray(ep["liveness"], dtype=np.int32),
            "draws": np.asarray(ep["draws"], dtype=np.float64),
            "agent_state": np.asarray(ep["agent_state"], dtype=np.float64),
        }
        if extra is not None:
            arrays["extra"] = np.asarray(ep["extra"], dtype=np.float64)
        out.append(arrays)
    return meta, out


def geometry_vectors(seed=0, count=200):
    """Known-answer vectors for the geometry bridge, computed BY THE REFERENCE
    (library/geometry.py:74-87 over the exact stand-in): random oriented boxes vs
    boxes -> intersects / contains / percentage_intersects, plus bounding_box corners
    (bodies.py:116-117) and stopping_zones (bodies.py:122-135)."""
    mods = refload.load()
    geometry, bodies = mods["geometry"], mods["bodies"]
    from examples.constants import car_constants, pedestrian_constants
    rs = np.random.RandomState(seed)
    rows = []
    for k in range(count):
        la, wa, lb, wb = rs.uniform(5, 80, size=4)
        xa, ya = rs.uniform(-50, 50, size=2)
        ta = rs.choice([0.0, math.pi, math.pi / 2, -math.pi / 2, rs.uniform(-math.pi, math.pi)])
        mode = k % 4
        if mode == 0:  # generic
            xb, yb = rs.uniform(-80, 80, size=2)
            tb = rs.uniform(-math.pi, math.pi)
        elif mode == 1:  # axis-aligned, exactly touching edges
            ta, tb = 0.0, 0.0
            xb, yb = xa + (la + lb) / 2, ya + rs.uniform(-5, 5)
        elif mode == 2:  # containment candidates
            lb, wb = la * rs.uniform(0.2, 1.2), wa * rs.uniform(0.2, 1.2)
            xb, yb = xa + rs.uniform(-3, 3), ya + rs.uniform(-3, 3)
            tb = ta if rs.uniform() < 0.5 else rs.uniform(-math.pi, math.pi)
        else:  # near misses
            tb = rs.uniform(-math.pi, math.pi)
            r = (math.hypot(la, wa) + math.hypot(lb, wb)) / 2 * rs.uniform(0.6, 1.0)
            phi = rs.uniform(-math.pi, math.pi)
            xb, yb = xa + r * math.cos(phi), ya + r * math.sin(phi)
        A = geometry.make_rectangle(la, wa).transform(ta, geometry.Point(xa, ya))
        B = geometry.make_rectangle(lb, wb).transform(tb, geometry.Point(xb, yb))
        rows.append([la, wa, xa, ya, ta, lb, wb, xb, yb, tb,
                     float(A.intersects(B)), float(B.contains(A)), float(A.percentage_intersects(B))]
                    + [c for p in A for c in p] + [c for p in B for c in p])
    zones = []
    for k in range(count // 4):
        consts = car_constants if k % 2 == 0 else pedestrian_constants
        x, y = rs.uniform(-100, 1600), rs.uniform(-100, 100)
        v = rs.uniform(0, consts.max_velocity) if k % 7 else 0.0
        th = rs.choice([0.0, math.pi, rs.uniform(-math.pi, math.pi)])
        body = bodies.Car(bodies.DynamicBodyState(geometry.Point(x, y), v, th), consts)
        bz, rz = body.stopping_zones()
        flat = [NAN] * 16 if bz is None else [c for p in bz for c in p] + [c for p in rz for c in p]
        zones.append([x, y, v, th, consts.length, consts.width, consts.min_throttle] + flat)
    return np.asarray(rows, dtype=np.float64), np.asarray(zones, dtype=np.float64)
