"""Agent plugin surface (reference examples/agents/): host-side agent classes with the
`reset / choose_action(state, action_space, info) / process_feedback(previous_state, action, state, reward)`
protocol.  They drive the single-environment compat view; in batched runs the same policies execute on the
device (csrc/agents.cuh) and `device_spec()` tells the engine which one to use."""
