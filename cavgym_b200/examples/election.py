"""Arbitration between ElectionAgent testers (reference examples/election.py:4-57): at most one pedestrian crosses at a
time.  `result(previous_state, joint_action)` is called by Simulation.run between choose_action and env.step
(simulation.py:73-74) and returns the joint action that is actually executed.

Rules, as the reference applies them each step:
  * while the current winner is still crossing, every other ElectionAgent is reset (its crossing plan is dropped) and its
    steering action zeroed — its throttle action stands;
  * otherwise the voters of this step (agents that just asked to cross) are ranked by distance to the ego, the closest wins
    (ties: `env.np_random.choice`, which consumes one draw of the env's shared RandomState even for a single candidate);
  * the winner is kept until its `crossing` flag drops.  The object outlives episodes, and so does `active_player`
    (it is only cleared on the first step of the next episode, once the agents' reset has dropped the flag).
Player 0 is the ego and never in the electorate, which is why the reference can test `if self.active_player`.
"""
from .agents.pedestrian import ElectionAgent


class Election:
    def __init__(self, env, agents):
        assert len(env.bodies) == len(agents)
        self.env, self.agents = env, agents
        self.electorate = [i for i, agent in enumerate(agents) if isinstance(agent, ElectionAgent)]
        self.active_player = None
        self.previous_joint_action = None

    def focal_joint_action(self, joint_action_vote, focal_player):
        assert focal_player in self.electorate
        for i in self.electorate:
            if i != focal_player:
                self.agents[i].reset()
                joint_action_vote[i] = [joint_action_vote[i][0], 0.0]
        return joint_action_vote

    def _closest_voters(self):
        ego = self.env.ego.state.position
        ranked = [(self.env.bodies[i].state.position.distance(ego), i) for i in self.electorate if self.agents[i].voting]
        if not ranked:
            return []
        nearest = min(distance for distance, _ in ranked)
        return [i for distance, i in ranked if distance == nearest]

    def result(self, previous_state, joint_action_vote):
        assert len(self.agents) == len(previous_state) == len(joint_action_vote)
        if self.active_player and not self.agents[self.active_player].crossing:
            self.active_player = None
        if not self.active_player:
            tied = self._closest_voters()
            winner = self.env.np_random.choice(tied) if tied else None
            if not winner:
                return joint_action_vote
            self.active_player = winner
        return self.focal_joint_action(joint_action_vote, self.active_player)
