"""``src.agent`` (agent/__init__.py:1-54: the three navigation agents, ``build_agent`` and the ``Speaker``)."""
from ..agent import build_agent, EnvDropAgent, FollowerAgent, SelfMonitorAgent, Speaker  # noqa: F401
