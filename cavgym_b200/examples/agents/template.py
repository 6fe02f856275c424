"""`examples.agents.template` of the reference (template.py:8-62) — the classes live in `_protocol`."""
from ._protocol import Agent, NoopAgent, RandomAgent  # noqa: F401
