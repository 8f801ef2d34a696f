"""``src.agent`` (agent/__init__.py:11-54)."""
from ..agent import build_agent, EnvDropAgent, FollowerAgent, SelfMonitorAgent  # noqa: F401
