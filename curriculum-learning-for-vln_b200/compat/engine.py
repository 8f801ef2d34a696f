"""``src.engine`` (engine/__init__.py:6-17, curriculum.py) — the trainers main.py:96-125 instantiates."""
from ..engine import ClassicTrainer, NaiveCurriculum, SelfPacedCurriculum, Evaluation, evaluate  # noqa: F401


def check_the_code(cfg, device, tok, valid_env):
    """trainer.py:27-39: follow the teacher through every validation episode and score it (expects SR ~ 1)."""
    import torch
    from ..agent import build_agent
    out = {}
    for key, env in valid_env.items():
        agent = build_agent(cfg, tok, device)
        agent.env = env
        agent.eval()
        with torch.no_grad():
            agent.test(iters=None, feedback="teacher")
        out[key] = evaluate(env, agent.get_results())
    return out
