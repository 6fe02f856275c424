"""Agent base classes (reference examples/agents/template.py:8-62)."""
import numpy as np

from ...scenario import AgentSpec


class Agent:
    def __init__(self, index, **kwargs):
        super().__init__(**kwargs)
        self.index = index

    def reset(self):
        raise NotImplementedError

    def choose_action(self, state, action_space, info=None):
        raise NotImplementedError

    def process_feedback(self, previous_state, action, state, reward):
        raise NotImplementedError

    def device_spec(self):
        """AgentSpec of the on-device equivalent of this agent, or None if it only exists on the host."""
        return None


class NoopAgent(Agent):
    def __init__(self, noop_action, **kwargs):
        super().__init__(**kwargs)
        self.noop_action = noop_action

    def reset(self):
        pass

    def choose_action(self, state, action_space, info=None):
        return self.noop_action

    def process_feedback(self, previous_state, action, state, reward):
        pass

    def device_spec(self):
        return AgentSpec("noop")


class RandomAgent(NoopAgent):
    """With probability epsilon per step draws a fresh action from the action space and holds it."""

    def __init__(self, epsilon, np_random=None, **kwargs):
        super().__init__(**kwargs)
        self.epsilon = epsilon
        self.np_random = np_random if np_random is not None else np.random.RandomState()
        self.action = self.noop_action

    def reset(self):
        self.action = self.noop_action

    def epsilon_valid(self):
        return self.np_random.uniform(0.0, 1.0) < self.epsilon

    def choose_action(self, state, action_space, info=None):
        if self.epsilon_valid():
            sample = action_space.sample()
            self.action = list(sample) if isinstance(sample, np.ndarray) else sample
        return self.action

    def device_spec(self):
        return AgentSpec("random", epsilon=self.epsilon)
