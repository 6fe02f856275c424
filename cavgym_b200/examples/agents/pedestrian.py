"""Pedestrian testers (reference examples/agents/pedestrian.py:14-116): the road-crossing state machine and its
random / proximity triggers.  Host classes for the single-environment compat view; csrc/agents.cuh
(choose_crossing_action, crossing_feedback) runs the same machine on the device.  ElectionAgent (arbitrated by
examples/election.py) is host-side; the reference's Q-learning tester (pedestrian.py:119-251) cannot run — its first
process_feedback multiplies the config's LinSpace object by a float (pedestrian.py:128, 247: TypeError; reproduced by
oracle/gen_learning_golden.py) — so it is not mirrored and Config.setup says so."""
import math

from ...library.geometry import Point
from ...scenario import AgentSpec
from .dynamic_body import TARGET_ERROR, make_body_state, make_steering_action
from .template import NoopAgent, RandomAgent


class CrossingAgent(NoopAgent):
    def __init__(self, body, time_resolution, road_centre, **kwargs):
        super().__init__(noop_action=body.noop_action, **kwargs)
        self.body, self.time_resolution, self.road_centre = body, time_resolution, road_centre
        self.reset()

    def reset(self):
        self.initial_distance = self.waypoint = self.target_orientation = self.prior_orientation = None

    def choose_action(self, state, action_space, info=None):
        raise NotImplementedError

    def idle(self):
        return self.waypoint is None and self.target_orientation is None

    def process_feedback(self, previous_state, action, state, reward):
        body_state = make_body_state(state, self.index)
        if self.waypoint is not None and body_state.position.distance(self.waypoint) < 1:
            self.waypoint, self.target_orientation, self.prior_orientation = None, self.prior_orientation, None
        if self.target_orientation is not None:
            diff = self.target_orientation - body_state.orientation
            if abs(math.atan2(math.sin(diff), math.cos(diff))) < TARGET_ERROR:
                self.target_orientation = None

    def choose_crossing_action(self, state, condition):
        body_state = make_body_state(state, self.index)
        initiated = self.idle() and bool(condition)
        if initiated:
            here = body_state.position
            foot = self.road_centre.closest_point_from(here)
            bearing = math.atan2(foot.y - here.y, foot.x - here.x)
            if self.initial_distance is None:
                self.initial_distance = here.distance(foot)
            self.waypoint = Point(foot.x + self.initial_distance * math.cos(bearing), foot.y + self.initial_distance * math.sin(bearing))
            self.target_orientation = math.atan2(self.waypoint.y - here.y, self.waypoint.x - here.x)
            self.prior_orientation = body_state.orientation
        steering = make_steering_action(body_state, self.body.constants, self.time_resolution, self.target_orientation, self.noop_action)
        return [self.noop_action[0], steering], initiated


class RandomConstrainedAgent(CrossingAgent, RandomAgent):
    def choose_action(self, state, action_space, info=None):
        return self.choose_crossing_action(state, self.epsilon_valid())[0]

    def device_spec(self):
        return AgentSpec("random-constrained", epsilon=self.epsilon)


class ProximityAgent(CrossingAgent):
    def __init__(self, distance_threshold, **kwargs):
        super().__init__(**kwargs)
        self.distance_threshold = distance_threshold

    def proximity_trigger(self, state):
        return Point(*state[self.index][:2]).distance(Point(*state[0][:2])) < self.distance_threshold

    def choose_action(self, state, action_space, info=None):
        return self.choose_crossing_action(state, self.proximity_trigger(state))[0]

    def device_spec(self):
        return AgentSpec("proximity", threshold=self.distance_threshold)


class ElectionAgent(ProximityAgent):
    """A ProximityAgent that only ASKS to cross (reference pedestrian.py:94-116): `voting` is raised on the step it would
    start a crossing, `crossing` from then until the crossing state machine is idle again.  Election (examples/election.py)
    lets one voter at a time carry its crossing out and resets the others."""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.voting = self.crossing = False

    def reset(self):
        super().reset()
        self.voting = self.crossing = False

    def choose_action(self, state, action_space, info=None):
        action, self.voting = self.choose_crossing_action(state, self.proximity_trigger(state))
        self.crossing = self.crossing or self.voting
        return action

    def process_feedback(self, previous_state, action, state, reward):
        super().process_feedback(previous_state, action, state, reward)
        if self.idle():
            self.crossing = False

    def device_spec(self):
        return None   # arbitration across agents happens on the host (examples/election.py)

