"""Container-only: oracle/port_beam.py against the UNMODIFIED reference agents' ``_dijkstra`` (src/agent/base.py:183-397)
for EnvDrop, Follower and Self-Monitor on the FakeSim world: the same K best paths per episode — trajectories, actions, listener scores,
visual features — and the same navigation path (``dijk_path``)."""
import random
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/repo")
import clvln_b200  # noqa: E402,F401
from clvln_b200.environ import make_items, make_world  # noqa: E402
from oracle import port_beam as PB, port_env as PE, port_rollout as PR, ref_harness as H  # noqa: E402

w = make_world(n_scans=3, seed=1)
items = make_items(w, 40, seed=1)
src = H.install(w, {"train": items})
import src.agent as agent_mod  # noqa: E402
import src.environ as environ  # noqa: E402

tok = H.StubTokenizer(items)
fs = H.feature_store(w)
view = PE.WorldView(w)
dev = torch.device("cpu")
for kind in ("ENVDROP", "FOLLOWER", "MONITOR"):
    random.seed(2020)
    torch.manual_seed(2020)
    renv = environ.R2RBatch(fs, batch_size=6, splits=["train"], tokenizer=tok)
    H.warm_candidate_buffer(renv)
    cfg = H.model_cfg(kind)
    if kind == "ENVDROP":
        ag = agent_mod.EnvDropAgent(cfg, 80, "/tmp", dev, renv, tok, episode_len=12)
    elif kind == "FOLLOWER":
        ag = agent_mod.FollowerAgent(cfg, "/tmp", dev, renv, tok, episode_len=10)
    else:
        ag = agent_mod.SelfMonitorAgent(cfg, 80, "/tmp", dev, renv, tok, episode_len=10)
        ag.reset_loss()
    ag.env = renv
    ag.eval()
    st = random.getstate()
    mods = [ag.encoder, ag.decoder] + ([ag.critic] if kind == "ENVDROP" else [])
    sds = [{k: v.detach().clone() for k, v in m.state_dict().items()} for m in mods]
    pag = PR.Agent(kind, sds[0], sds[1], sds[2] if kind == "ENVDROP" else None, hidden=cfg.HIDDEN_SIZE,
                   bidirectional=cfg.ENC_BIDIRECTION, enc_layers=cfg.ENC_LAYERS, episode_len=ag.episode_len)
    random.seed(2020)
    penv = PE.R2RBatchPort(view, items, batch_size=6)
    random.setstate(st)
    for K in (3, 5):
        with torch.no_grad():
            ref = ag._dijkstra(K)
            got = PB.dijkstra(pag, penv, K, full_length=(kind == "MONITOR"))   # monitor.py:68-87: always the full 80 tokens
        assert [r["instr_id"] for r in ref] == [r["instr_id"] for r in got]
        n_paths = 0
        for r, g in zip(ref, got):
            assert r["dijk_path"] == g["dijk_path"], (r["dijk_path"], g["dijk_path"])
            assert len(r["paths"]) == len(g["paths"])
            rp = sorted(r["paths"], key=lambda p: (tuple(p["action"]), tuple(x[0] for x in p["trajectory"])))
            gp = sorted(g["paths"], key=lambda p: (tuple(p["action"]), tuple(x[0] for x in p["trajectory"])))
            for a, b in zip(rp, gp):
                assert a["trajectory"] == b["trajectory"], (a["trajectory"], b["trajectory"])
                assert a["action"] == b["action"] and a["listener_actions"] == b["listener_actions"]
                assert np.allclose(a["listener_scores"], b["listener_scores"], rtol=1e-5, atol=1e-6)
                for (f1, c1), (f2, c2) in zip(a["visual_feature"], b["visual_feature"]):
                    assert torch.equal(f1, f2) and torch.equal(c1, c2)
                n_paths += 1
        print(f"{kind} K={K}: {n_paths} paths over {len(ref)} episodes identical")
print("beam port OK")
