"""Agents (reference: src/agent/__init__.py:11-54 build_agent and the three agent classes)."""
from .base import BaseAgent, RolloutState
from .envdrop import EnvDropAgent
from .follower import FollowerAgent
from .monitor import SelfMonitorAgent
from .speaker import Speaker


def build_agent(cfg, tokenizer, device, **kwargs):
    name = cfg.MODEL.NAME
    if name == "FOLLOWER":
        return FollowerAgent(model_cfg=cfg.MODEL.FOLLOWER, results_dir=cfg.OUTPUT.RESULT_DIR, device=device, env=None,
                             tokenizer=tokenizer, episode_len=cfg.AGENT.MAX_EPISODE_LEN)
    if name == "SELF-MONITOR":
        return SelfMonitorAgent(model_cfg=cfg.MODEL.MONITOR, max_enc_len=cfg.DATA.MAX_ENC_LEN,
                                results_dir=cfg.OUTPUT.RESULT_DIR, device=device, env=None, tokenizer=tokenizer,
                                episode_len=cfg.AGENT.MAX_EPISODE_LEN)
    if name == "ENVDROP":
        return EnvDropAgent(model_cfg=cfg.MODEL.ENVDROP, max_enc_len=cfg.DATA.MAX_ENC_LEN,
                            results_dir=cfg.OUTPUT.RESULT_DIR, device=device, env=None, tokenizer=tokenizer,
                            episode_len=cfg.AGENT.MAX_EPISODE_LEN)
    raise NotImplementedError(name)


__all__ = ["BaseAgent", "RolloutState", "EnvDropAgent", "FollowerAgent", "SelfMonitorAgent", "Speaker", "build_agent"]
