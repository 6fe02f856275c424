"""The learning ego (reference examples/agents/ego.py:16-145): linear Q-learning over look-ahead features.

  QLearningEgoAgent          the reference's agent for the single-environment compat view: one Python object, one weight
                             table, one TD(0) update per step.  Draw for draw and operation for operation the reference's
                             arithmetic, so a run through `Simulation` reproduces the reference's actions and weights
                             (tests/test_learning_agents.py, fixtures recorded from the unmodified reference).
  BatchedQLearningEgoAgent   the same learner on the tensor API of BatchedCAVEnv.  Features, Q values, the greedy choice and
                             the TD targets are torch expressions over [N] on the device the engine steps on (no host round
                             trip per step).  Two forms: `shared=True` — N environments feed ONE weight table, the N
                             transitions of a batch step applied as one averaged update (with N = 1 and the same draws it is
                             the reference's update sequence, checked against the host class); `shared=False` — every
                             environment is its OWN learner (own table, own alpha / gamma / epsilon): N independent runs of
                             the reference's agent side by side, which is what experiments.py's grid is (cavgym_b200/experiments.py).

Features (ego.py:96-145), per opponent i, after a two-step no-steering look-ahead of the ego under the candidate throttle
and of the opponent under zero throttle — both clamped with the EGO's velocity limits, as the reference does:
distance_x, distance_y, distance, relative_angle, heading; each normalised into [0, 1] by its bounds.
"""
import math

import numpy as np

from ... import reporting
from ...library import geometry
from ...library.geometry import Point
from .dynamic_body import make_body_state
from .template import RandomAgent

FEATURES = ("distance_x", "distance_y", "distance", "relative_angle", "heading")   # the order ego.py:39-48 enables them in
LOOKAHEAD_STEPS = 2


def feature_bounds(feature_config, width, height):
    """{feature: (low, high)} for the enabled features, in FEATURES order (ego.py:38-48)."""
    bounds = {"distance_x": (-float(width), float(width)), "distance_y": (-float(height), float(height)),
              "distance": (0.0, math.sqrt((width ** 2) + (height ** 2))), "relative_angle": (0.0, math.pi), "heading": (0.0, math.pi)}
    return {name: bounds[name] for name in FEATURES if getattr(feature_config, name)}


def _unit(value, low, high):
    return 0.0 if value < low else 1.0 if value > high else (value - low) / (high - low)


class QLearningEgoAgent(RandomAgent):
    def __init__(self, q_learning_config, body, time_resolution, num_opponents, num_actions, width, height, **kwargs):
        super().__init__(noop_action=body.noop_action, epsilon=q_learning_config.epsilon, **kwargs)
        schedule = q_learning_config.alpha
        self.target_alpha = schedule.stop
        self.alphas = iter(np.linspace(start=schedule.start, stop=schedule.stop, num=schedule.num_steps, endpoint=True))
        self.alpha = next(self.alphas, self.target_alpha)
        self.gamma = q_learning_config.gamma
        self.feature_config = q_learning_config.features
        self.body, self.time_resolution = body, time_resolution
        self.opponent_indexes = list(range(1, num_opponents + 1))
        k = body.constants
        self.available_actions = [[throttle, self.noop_action[1]]
                                  for throttle in np.linspace(start=k.min_throttle, stop=k.max_throttle, num=num_actions, endpoint=True)]
        self.feature_bounds = feature_bounds(self.feature_config, width, height)
        self.feature_weights = {index: dict.fromkeys(self.feature_bounds, 0.0) for index in self.opponent_indexes}
        self.log_file = reporting.get_agent_file_logger(q_learning_config.log) if q_learning_config.log is not None else None
        if self.log_file:
            self.enabled_features = {index: sorted(self.feature_bounds) for index in self.opponent_indexes}
            self.log_file.info(",".join(f"{name}{index}" for index, names in self.enabled_features.items() for name in names))

    def reset(self):
        pass   # the weights survive episodes; only RandomAgent's held action would be reset, and this agent holds none

    # ---- features --------------------------------------------------------------------------------------------------
    def _ahead(self, body_state, throttle):
        """LOOKAHEAD_STEPS straight-line steps of DynamicBody.step at constant throttle (ego.py:100-116)."""
        k, dt = self.body.constants, self.time_resolution
        x, y = body_state.position
        velocity, heading = body_state.velocity, body_state.orientation
        for _ in range(LOOKAHEAD_STEPS):
            travelled = velocity * dt
            x, y = x + travelled * math.cos(heading), y + travelled * math.sin(heading)
            velocity = max(k.min_velocity, min(k.max_velocity, velocity + (throttle * dt)))
        return Point(x, y), heading

    def features_opponent(self, state, action, opponent_index):
        me, my_heading = self._ahead(make_body_state(state, self.index), action[0])
        other, other_heading = self._ahead(make_body_state(state, opponent_index), 0.0)
        raw = {"distance_x": lambda: me.distance_x(other), "distance_y": lambda: me.distance_y(other), "distance": lambda: me.distance(other),
               "relative_angle": lambda: abs(geometry.normalise_angle(geometry.Line(start=me, end=other).orientation() - my_heading)),
               "heading": lambda: abs(geometry.normalise_angle(geometry.Line(start=other, end=me).orientation() - other_heading))}
        return {name: _unit(raw[name](), *bounds) for name, bounds in self.feature_bounds.items()}

    def features(self, state, action):
        return {index: self.features_opponent(state, action, index) for index in self.opponent_indexes}

    def q_value(self, state, action):
        return sum(value * self.feature_weights[index][name]
                   for index, values in self.features(state, action).items() for name, value in values.items())

    # ---- policy and update -----------------------------------------------------------------------------------------
    def _pick(self, candidates):
        return candidates[0] if len(candidates) == 1 else candidates[self.np_random.choice(range(len(candidates)))]

    def choose_action(self, state, action_space, info=None):
        if self.epsilon_valid():
            return self.available_actions[self.np_random.choice(range(len(self.available_actions)))]
        best, best_q = [], -math.inf
        for action in self.available_actions:   # every action tied for the largest Q value, in action order
            q = self.q_value(state, action)
            if q > best_q:
                best, best_q = [action], q
            elif q == best_q:
                best.append(action)
        assert best, "no best action(s) found"
        return self._pick(best)

    def process_feedback(self, previous_state, action, state, reward):
        target = reward + self.gamma * max(self.q_value(state, candidate) for candidate in self.available_actions)
        difference = target - self.q_value(previous_state, action)
        for index, values in self.features(previous_state, action).items():
            weights = self.feature_weights[index]
            for name, value in values.items():
                weights[name] = weights[name] + self.alpha * difference * value
        if self.log_file:
            self.log_file.info(",".join(str(self.feature_weights[index][name]) for index, names in self.enabled_features.items() for name in names))
        self.alpha = next(self.alphas, self.target_alpha)

    def device_spec(self):
        return None   # learns on the host; BatchedQLearningEgoAgent is its tensor-API form


class BatchedQLearningEgoAgent:
    """QLearningEgoAgent over the N environments of a BatchedCAVEnv, every tensor on the engine's device.

    choose_action(state [M,4,N]) -> (action index [N] int64, ego action rows [2,N]); process_feedback(previous_state,
    action index, state, ego reward [N], live [N] bool) applies  w += alpha * mean_e(difference_e * features_e)  over the
    live environments (shared table, `weights` [opponents, features]) or  w_e += alpha_e * difference_e * features_e  per
    environment (independent learners, `weights` [N, opponents, features]) and advances the alpha schedule once."""

    def __init__(self, q_learning_config, ego_constants, time_resolution, num_opponents, width, height, num_envs, device,
                 num_actions=5, dtype=None, seed=0, shared=True, alpha=None, gamma=None, epsilon=None):
        """`alpha` = (start, stop, num_steps), `gamma`, `epsilon`: per-environment overrides of the config's values, each a
        scalar or a length-N sequence (independent learners of a hyper-parameter grid)."""
        import torch
        self.torch = torch
        self.device, self.dtype = device, dtype or torch.float64
        self.shared = bool(shared)
        self.num_envs, self.num_opponents = int(num_envs), int(num_opponents)

        def per_env(value, default):
            value = default if value is None else value
            if np.ndim(value) == 0:
                return float(value)
            return torch.as_tensor(np.asarray(value, dtype=np.float64), device=device).reshape(self.num_envs)

        schedule = q_learning_config.alpha
        start, stop, num_steps = alpha if alpha is not None else (schedule.start, schedule.stop, schedule.num_steps)
        self.epsilon, self.gamma = per_env(epsilon, q_learning_config.epsilon), per_env(gamma, q_learning_config.gamma)
        self._alpha_start, self._alpha_stop, self._alpha_steps = per_env(start, None), per_env(stop, None), per_env(num_steps, None)
        if all(np.ndim(v) == 0 for v in (start, stop, num_steps)):   # the reference's own table of values, bit for bit
            self._alphas = np.linspace(start=start, stop=stop, num=int(num_steps), endpoint=True)
        else:
            self._alphas = None
        self._alpha_at, self.target_alpha = 0, stop
        self.k, self.dt = ego_constants, float(time_resolution)
        self.bounds = feature_bounds(q_learning_config.features, width, height)
        self.names = list(self.bounds)
        self.throttles = torch.linspace(ego_constants.min_throttle, ego_constants.max_throttle, num_actions, dtype=torch.float64, device=device)
        table = (self.num_opponents, len(self.names))
        self.weights = torch.zeros(table if self.shared else (self.num_envs,) + table, dtype=torch.float64, device=device)
        self.generator = torch.Generator(device=device)
        self.generator.manual_seed(int(seed))

    @property
    def alpha(self):
        """Learning rate of the coming update: a float, or [N] when the schedules differ per environment."""
        if self._alphas is not None:
            return float(self._alphas[self._alpha_at]) if self._alpha_at < len(self._alphas) else float(self.target_alpha)
        torch = self.torch
        steps = torch.as_tensor(self._alpha_steps, dtype=torch.float64, device=self.device)
        at = torch.minimum(torch.full_like(steps, float(self._alpha_at)), steps - 1.0)
        return self._alpha_start + (self._alpha_stop - self._alpha_start) * at / (steps - 1.0)   # np.linspace, element by element

    def _ahead(self, rows, throttle):
        """rows [4, ...] -> (x, y, heading) after the look-ahead; `throttle` broadcasts against rows[0]."""
        torch = self.torch
        x, y, velocity, heading = rows[0], rows[1], rows[2], rows[3]
        cos, sin = torch.cos(heading), torch.sin(heading)
        for _ in range(LOOKAHEAD_STEPS):
            travelled = velocity * self.dt
            x, y = x + travelled * cos, y + travelled * sin
            velocity = torch.clamp(velocity + throttle * self.dt, self.k.min_velocity, self.k.max_velocity)
        return x, y, heading

    @staticmethod
    def _wrap(torch, angle):   # geometry.normalise_angle for |angle| < 3 pi
        angle = torch.where(angle <= -math.pi, angle + 2 * math.pi, angle)
        return torch.where(angle > math.pi, angle - 2 * math.pi, angle)

    def features(self, state):
        """[actions, opponents, features, N] for every candidate throttle."""
        torch = self.torch
        state = state.to(torch.float64)
        ego = state[0].unsqueeze(1)                                    # [4, 1, N]
        mx, my, mh = self._ahead(ego, self.throttles.view(-1, 1))      # [A, N]
        ox, oy, oh = self._ahead(state[1:1 + self.num_opponents].permute(1, 0, 2), 0.0)   # [O, N]
        dx, dy = ox.unsqueeze(0) - mx.unsqueeze(1), oy.unsqueeze(0) - my.unsqueeze(1)     # [A, O, N]
        raw = {"distance_x": lambda: dx.abs(), "distance_y": lambda: dy.abs(), "distance": lambda: torch.sqrt(dy * dy + dx * dx),
               "relative_angle": lambda: self._wrap(torch, torch.atan2(dy, dx) - mh.unsqueeze(1)).abs(),
               "heading": lambda: self._wrap(torch, torch.atan2(-dy, -dx) - oh.unsqueeze(0)).abs()}
        columns = []
        for name in self.names:
            low, high = self.bounds[name]
            columns.append(torch.clamp((raw[name]() - low) / (high - low), 0.0, 1.0))
        return torch.stack(columns, dim=2)

    def q_values(self, state):
        """([A, N] Q values, features)"""
        features = self.features(state)
        weights = self.weights.view(1, self.num_opponents, -1, 1) if self.shared else self.weights.permute(1, 2, 0).unsqueeze(0)
        return (features * weights).sum(dim=(1, 2)), features

    def choose_action(self, state, u_explore=None, u_pick=None):
        """epsilon-greedy per environment; ties for the largest Q value are broken uniformly, in action order, like
        `np_random.choice(best_actions)`.  The two uniforms per environment may be supplied (tests replay recorded draws)."""
        torch = self.torch
        n, actions = self.num_envs, self.throttles.numel()
        if u_explore is None:
            u_explore = torch.rand(n, dtype=torch.float64, device=self.device, generator=self.generator)
        if u_pick is None:
            u_pick = torch.rand(n, dtype=torch.float64, device=self.device, generator=self.generator)
        q, _ = self.q_values(state)
        tied = q == q.max(dim=0, keepdim=True).values                                   # [A, N]
        explore = u_explore < self.epsilon
        candidates = torch.where(explore.unsqueeze(0), torch.ones_like(tied), tied)
        count = candidates.sum(dim=0)
        wanted = torch.clamp((u_pick * count).floor().long(), max=actions - 1).minimum(count - 1)   # k-th candidate, k from 0
        # the wanted-th candidate in action order sits after exactly `wanted` candidates: count the positions whose running
        # candidate count is still <= wanted (the running count is non-decreasing along the action axis)
        index = (candidates.long().cumsum(dim=0) <= wanted.unsqueeze(0)).sum(dim=0).clamp(max=actions - 1)
        rows = torch.stack([self.throttles[index], torch.zeros(n, dtype=torch.float64, device=self.device)])
        return index, rows.to(self.dtype)

    def process_feedback(self, previous_state, action_index, state, reward, live=None):
        torch = self.torch
        q_next, _ = self.q_values(state)
        q_prev, features_prev = self.q_values(previous_state)
        env = torch.arange(self.num_envs, device=self.device)
        chosen = features_prev[action_index, :, :, env]                                  # [N, O, F]
        difference = (reward.to(torch.float64) + self.gamma * q_next.max(dim=0).values) - q_prev[action_index, env]
        if live is None:
            live = torch.ones(self.num_envs, dtype=torch.bool, device=self.device)
        weight = live.to(torch.float64)
        if not self.shared:   # N independent learners: each env applies its own update, as the reference's agent does
            alpha = self.alpha
            alpha = alpha.view(-1, 1, 1) if torch.is_tensor(alpha) else alpha
            self.weights = self.weights + alpha * (difference * weight).view(-1, 1, 1) * chosen
            self._alpha_at += 1
            return
        packed = torch.cat([((difference * weight).view(-1, 1, 1) * chosen).sum(dim=0).reshape(-1), weight.sum().view(1)])
        if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
            torch.distributed.all_reduce(packed)   # env-sharded ranks learn ONE table: sum of the updates and of the live counts
        self.weights = self.weights + self.alpha * packed[:-1].view_as(self.weights) / packed[-1].clamp(min=1.0)
        self._alpha_at += 1

    def feature_weights(self):
        """{opponent index: {feature: weight}} like the host agent's table."""
        table = (self.weights if self.shared else self.weights.mean(dim=0)).cpu().tolist()   # independent learners: their mean
        return {i + 1: dict(zip(self.names, table[i])) for i in range(self.num_opponents)}
