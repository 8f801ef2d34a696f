from .config import CfgNode, get_cfg_defaults, agent_cfg, StubTokenizer, length2mask  # noqa: F401
