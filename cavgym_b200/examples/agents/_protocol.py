"""Agent protocol and the two stock policies that need no knowledge of the body they drive (reference
examples/agents/template.py:8-62): `reset()`, `choose_action(state, action_space, info)`,
`process_feedback(previous_state, action, state, reward)`.  These host-side classes drive the single-environment compat
view; `device_spec()` names the on-device twin (csrc/agents.cuh) that batched runs use instead."""
import numpy as np

from ...scenario import AgentSpec


class Agent:
    """What Simulation.run expects of an agent; `index` is the position of its body in env.bodies."""

    def __init__(self, index, **kwargs):
        super().__init__(**kwargs)
        self.index = index

    def _abstract(self, *_args, **_kwargs):
        raise NotImplementedError

    reset = choose_action = process_feedback = _abstract

    def device_spec(self):
        """AgentSpec of the on-device equivalent of this agent, or None if it only exists on the host."""
        return None


class NoopAgent(Agent):
    """Always the body's noop action; learns nothing."""

    def __init__(self, noop_action, **kwargs):
        super().__init__(**kwargs)
        self.noop_action = noop_action

    def reset(self):
        return None

    def choose_action(self, state, action_space, info=None):
        return self.noop_action

    def process_feedback(self, previous_state, action, state, reward):
        return None

    def device_spec(self):
        return AgentSpec("noop")


class RandomAgent(NoopAgent):
    """Holds an action; with probability epsilon per step replaces it by a fresh sample of the action space
    (one uniform draw for the test, then Box.sample / Discrete.sample on the shared RandomState)."""

    def __init__(self, epsilon, np_random=None, **kwargs):
        super().__init__(**kwargs)
        self.epsilon = epsilon
        # the reference's default is an unseeded generator made at import time (template.py:41)
        self.np_random = np.random.RandomState() if np_random is None else np_random
        self.action = self.noop_action

    def epsilon_valid(self):
        return self.np_random.uniform(0.0, 1.0) < self.epsilon

    def reset(self):
        self.action = self.noop_action

    def choose_action(self, state, action_space, info=None):
        if self.epsilon_valid():
            drawn = action_space.sample()
            self.action = drawn.tolist() if isinstance(drawn, np.ndarray) else drawn
        return self.action

    def device_spec(self):
        return AgentSpec("random", epsilon=self.epsilon)
