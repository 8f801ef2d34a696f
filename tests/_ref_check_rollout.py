import sys, random, copy, numpy as np, torch
sys.path.insert(0, "/root/repo")
import clvln_b200
from clvln_b200.environ import make_world, make_items
from oracle import ref_harness as H, ref_loader, port_env as PE, port_rollout as PR, port_modules as P
w = make_world(n_scans=3, seed=1)
items = make_items(w, 40, seed=1)
src = H.install(w, {"train": items})
import src.environ as environ, src.agent as agent_mod
tok = H.StubTokenizer(items); fs = H.feature_store(w)
view = PE.WorldView(w)
dev = torch.device("cpu")
def grads(params): return torch.cat([p.grad.flatten() if p.grad is not None else torch.zeros_like(p).flatten() for p in params])
def cmp(a, b, name, tol=1e-5):
    d = (a-b).abs().max().item(); rel = d / max(b.abs().max().item(), 1e-12); print(f"  {name}: abs {d:.3e} rel {rel:.3e}"); assert rel < tol, name

for mode in ("eval", "train"):
  for kind in ("ENVDROP", "FOLLOWER", "MONITOR"):
    print(kind, mode)
    random.seed(2020); torch.manual_seed(2020)
    renv = environ.R2RBatch(fs, batch_size=8, splits=["train"], tokenizer=tok); H.warm_candidate_buffer(renv)
    cfg = H.model_cfg(kind)
    if kind == "ENVDROP": ag = agent_mod.EnvDropAgent(cfg, 80, "/tmp", dev, renv, tok, episode_len=12)
    elif kind == "FOLLOWER": ag = agent_mod.FollowerAgent(cfg, "/tmp", dev, renv, tok, episode_len=10)
    else: ag = agent_mod.SelfMonitorAgent(cfg, 80, "/tmp", dev, renv, tok, episode_len=10); ag.reset_loss()
    ag.env = renv
    getattr(ag, mode)()
    st = random.getstate()
    # port agent with cloned weights
    sds = [ {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and "running" not in k and k != "position.pe") for k, v in m.state_dict().items()} for m in ([ag.encoder, ag.decoder] + ([ag.critic] if kind=="ENVDROP" else [])) ]
    pag = PR.Agent(kind, sds[0], sds[1], sds[2] if kind=="ENVDROP" else None, hidden=cfg.HIDDEN_SIZE, bidirectional=cfg.ENC_BIDIRECTION, enc_layers=cfg.ENC_LAYERS, episode_len=ag.episode_len)
    pag.training = (mode == "train")
    drop = P.Drop("torch") if mode == "train" else None
    # --- reference
    torch.manual_seed(7)
    if kind == "ENVDROP":
        t1 = ag.rollout(train_ml=True, train_rl=False, feedback="teacher"); ml = ag.loss["ml_loss"]
        t2 = ag.rollout(train_ml=False, train_rl=True, restart=True, feedback="sample"); rl = ag.loss["rl_loss"]
        loss = ml + rl
    elif kind == "FOLLOWER":
        t1 = ag.rollout(feedback="teacher"); loss = ag.ml_loss
        t2 = ag.rollout(feedback="sample", train_cl=True); loss = loss + ag.ml_loss.sum()
    else:
        t1 = ag.rollout(feedback="teacher"); loss = ag.ml_loss
        t2 = ag.rollout(feedback="sample", train_cl=True); loss = loss + ag.ml_loss.sum()
    loss.backward()
    rparams = [p for m in ([ag.encoder, ag.decoder] + ([ag.critic] if kind=="ENVDROP" else [])) for p in m.parameters()]
    gref = grads(rparams)
    # --- port
    random.setstate(st)
    random.seed(2020)
    penv = PE.R2RBatchPort(view, items, batch_size=8)
    assert [d["instr_id"] for d in penv.data] == [d["instr_id"] for d in renv.data]
    random.setstate(st)
    torch.manual_seed(7)
    if kind == "ENVDROP":
        p1, l1 = PR.rollout_envdrop(pag, penv, train_ml=True, train_rl=False, feedback="teacher", drop=drop)
        p2, l2 = PR.rollout_envdrop(pag, penv, train_ml=False, train_rl=True, restart=True, feedback="sample", drop=drop)
        ploss = l1["ml_loss"] + l2["rl_loss"]
    elif kind == "FOLLOWER":
        p1, l1 = PR.rollout_follower(pag, penv, feedback="teacher", drop=drop)
        p2, l2 = PR.rollout_follower(pag, penv, feedback="sample", train_cl=True, drop=drop); ploss = l1 + l2.sum()
    else:
        p1, l1, _ = PR.rollout_monitor(pag, penv, feedback="teacher", drop=drop)
        p2, l2, _ = PR.rollout_monitor(pag, penv, feedback="sample", train_cl=True, drop=drop); ploss = l1 + l2.sum()
    ploss.backward()
    assert t1 == p1 and t2 == p2, "trajectories differ"
    pparams = [v for sd in sds for k, v in sd.items() if v.requires_grad]
    assert len(pparams) == len(rparams)
    print("  loss", loss.item(), ploss.item(), "steps", len(t2[0]["path"]))
    cmp(ploss.detach(), loss.detach(), "loss"); cmp(grads(pparams), gref, "grads", 1e-4)
print("rollout port OK")
