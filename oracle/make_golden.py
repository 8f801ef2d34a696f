"""Generate tests/golden/*.pt from the REAL reference (container only: needs /root/reference).

TEST INFRASTRUCTURE.  The reference has no golden vectors for this path, so they are produced by
running its unmodified code here (through oracle/ref_loader.py + oracle/ref_harness.py) on seeded
inputs; the committed fixtures then pin oracle/port_*.py wherever /root/reference is absent
(tests/test_golden.py).  Two kinds:

  modules.pt   small-dimension instances of EncoderLSTM (3 layouts), EnvDropDecoder,
               AttnDecoderLSTM, MonitorDecoder (eval and BN-training), Critic: state_dict, inputs,
               outputs — a few hundred KB.
  speaker.pt   the real Speaker class (shipped sizes, eval mode) on the seeded world: path lengths, loss / accuracies,
               beam-search scores, greedy words, gradient norms.
  beam.pt      the real EnvDrop / Follower agents' _dijkstra (K = 3): paths, actions, listener scores, dijk_path.
  rollouts.pt  the three real agents (shipped model sizes) on a seeded synthetic world:
               teacher + forced-action rollouts in eval mode: per-step logits/targets, losses,
               per-parameter gradient norms, trajectories, and the minibatch order over a
               wrap-around.  World and weights are regenerated from seeds by the test.

usage: python -m oracle.make_golden [speaker|beam]
"""
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")


def module_cases():
    from oracle import ref_loader
    U, Pol = ref_loader.load_ref_models()
    out = {}
    torch.manual_seed(1234)
    B, L, C, H, IMG = 4, 9, 5, 32, 16
    F_ = IMG + 128
    lens = torch.tensor([9, 7, 4, 2])
    toks = torch.randint(4, 60, (B, L))
    for i, l in enumerate(lens):
        toks[i, l:] = 0
    for name, (E, Hh, bi, nl) in {"enc_bi1": (24, H, True, 1), "enc_bi2": (20, H, True, 2), "enc_uni": (24, H, False, 1)}.items():
        enc = U.EncoderLSTM(60, E, Hh, 0, 0.5, bi, nl).eval()
        ctx, h, c = enc(toks, lens)
        out[name] = dict(sd={k: v.clone() for k, v in enc.state_dict().items()}, cfg=(E, Hh, bi, nl), toks=toks, lens=lens,
                         ctx=ctx.detach(), h=h.detach(), c=c.detach())
    img, cand = torch.randn(B, 36, F_), torch.randn(B, C, F_)
    mask = torch.zeros(B, L, dtype=torch.bool)
    mask[1, 7:] = True
    mask[3, 2:] = True
    ctx = torch.randn(B, L, H)
    a128, ht, c0 = torch.randn(B, 128), torch.randn(B, H), torch.randn(B, H)
    dec = Pol.EnvDropDecoder(H, 0.5, 0.3, action_embed_size=8, feature_size=F_).eval()
    lo, (h1, c1), htl = dec(a128, img.clone(), cand.clone(), ht, ht, c0, ctx, mask)
    out["envdrop"] = dict(sd=dec.state_dict(), a=a128, img=img, cand=cand, ht=ht, c0=c0, ctx=ctx, mask=mask,
                          logit=lo.detach(), h1=h1.detach(), c1=c1.detach(), h_tilde=htl.detach())
    dec = Pol.AttnDecoderLSTM(H, 0.5, action_embed_size=F_, feature_size=F_).eval()
    ap = torch.randn(B, F_)
    lo, (h1, c1), (ac, av) = dec(img, ap, cand, ht, c0, ctx, mask)
    out["follower"] = dict(sd=dec.state_dict(), img=img, ap=ap, cand=cand, h0=ht, c0=c0, ctx=ctx, mask=mask,
                           logit=lo.detach(), h1=h1.detach(), c1=c1.detach(), alpha_c=ac.detach(), alpha_v=av.detach())
    for training in (False, True):
        dec = Pol.MonitorDecoder(H, 0.5, 12, [40], action_embed_size=F_, feature_size=F_)
        dec.train(training)
        for m in dec.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
        sd0 = {k: v.clone() for k, v in dec.state_dict().items()}
        ctx12 = torch.randn(B, 12, H)
        m12 = torch.zeros(B, 12, dtype=torch.bool)
        m12[:, 8:] = True
        cm = torch.zeros(B, C, dtype=torch.bool)
        cm[0, 3:] = True
        cm[2, 2:] = True
        (lo, pr), (h1, c1), (ca, va) = dec(None, ap, cand, ht, c0, ctx12, m12, cm)
        out[f"monitor_train{int(training)}"] = dict(
            sd=sd0, ap=ap, cand=cand, h0=ht, c0=c0, ctx=ctx12, mask=m12, cmask=cm, logit=lo.detach(), prog=pr.detach(),
            h1=h1.detach(), c1=c1.detach(), ctx_attn=ca.detach(), cand_attn=va.detach(),
            rm=dec.state_dict()["proj_navigable_mlp.mlp.0.running_mean"].clone(),
            rv=dec.state_dict()["proj_navigable_mlp.mlp.2.running_var"].clone())
    cr = Pol.Critic(H, 0.5).eval()
    out["critic"] = dict(sd=cr.state_dict(), x=ht, y=cr(ht).detach())
    return out


def rollout_cases():
    import clvln_b200  # noqa: F401
    from clvln_b200.environ import make_world, make_items
    from oracle import ref_harness as H
    w = make_world(n_scans=3, seed=1)
    items = make_items(w, 40, seed=1)
    src = H.install(w, {"train": items})
    import src.environ as environ
    import src.agent as agent_mod
    tok = H.StubTokenizer(items)
    fs = H.feature_store(w)
    dev = torch.device("cpu")
    out = {"world": dict(n_scans=3, seed=1, n_items=40, B=8)}
    # minibatch order across two wrap-arounds
    random.seed(2020)
    env = environ.R2RBatch(fs, batch_size=16, splits=["train"], tokenizer=tok)
    random.seed(1)
    order = []
    for _ in range(7):
        env._next_minibatch()
        order.append([it["instr_id"] for it in env.batch])
    out["order"] = order
    for kind in ("ENVDROP", "FOLLOWER", "MONITOR"):
        random.seed(2020)
        torch.manual_seed(2020)
        renv = environ.R2RBatch(fs, batch_size=8, splits=["train"], tokenizer=tok)
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
        mods = [ag.encoder, ag.decoder] + ([ag.critic] if kind == "ENVDROP" else [])
        torch.manual_seed(7)
        if kind == "ENVDROP":
            t1 = ag.rollout(train_ml=True, train_rl=False, feedback="teacher")
            l1 = ag.loss["ml_loss"]
            t2 = ag.rollout(train_ml=False, train_rl=True, restart=True, feedback="sample")
            l2 = ag.loss["rl_loss"]
            loss = l1 + l2
        else:
            t1 = ag.rollout(feedback="teacher")
            l1 = ag.ml_loss
            t2 = ag.rollout(feedback="sample", train_cl=True)
            l2 = ag.ml_loss
            loss = l1 + l2.sum()
        loss.backward()
        gn = [float(p.grad.norm()) if p.grad is not None else 0.0 for m in mods for p in m.parameters()]
        out[kind] = dict(traj1=t1, traj2=t2, l1=l1.detach(), l2=l2.detach(), grad_norms=gn,
                         w_checksum=[float(p.detach().double().sum()) for m in mods for p in m.parameters()])
    return out


def speaker_cases():
    """The real Speaker class (shipped sizes) on the seeded synthetic world, eval mode: path lengths, teacher-forcing loss /
    accuracies, the per-word scores of the beam-search entry point, greedy words, gradient norms of the eval-mode loss;
    world and weights are regenerated from the seeds by the tests (weight checksums guard the init order).  Driven through
    the key-renaming adapter of tests/_ref_check_speaker.py (the class reads observation keys its own env lacks)."""
    import clvln_b200  # noqa: F401
    from clvln_b200.environ import make_world, make_items
    from oracle import ref_harness as H, ref_loader
    w = make_world(n_scans=3, seed=1)
    items = make_items(w, 40, seed=1)
    H.install(w, {"train": items})
    import src.agent.speaker as ref_speaker
    import src.environ as environ
    torch.Tensor.cuda = lambda self, *a, **k: self
    ref_speaker.Variable = torch.autograd.Variable
    if not hasattr(np, "bool"):
        np.bool = bool

    class OldKeys:
        def __init__(self, r2r):
            self.r2r, self.env = r2r, r2r.env
            self.feature_size, self.batch_size = r2r.feature_size, r2r.batch_size

        def reset(self, **kw):
            return self._wrap(self.r2r.reset(**kw))

        def _get_obs(self):
            return self._wrap(self.r2r.observe())

        @staticmethod
        def _wrap(obs):
            return [dict(ob, viewpoint=ob["viewpointId"],
                         candidate=[dict(c, pointId=c["absViewIndex"], viewpointId=c["nextViewpointId"]) for c in ob["candidates"]])
                    for ob in obs]

    tok = H.StubTokenizer(items)
    cfg = ref_loader.AttrDict(RNN_DIM=512, DROPOUT=0.6, FEAT_DROPOUT=0.3, BI_DIRECTION=True, WEMB=256, LR=1e-4,
                              FAST_TRAIN=False, IGNORE_ID=-1, MAX_DECODE=24, LOAD_OPTIM=False)
    random.seed(2020)
    torch.manual_seed(2020)
    renv = environ.R2RBatch(H.feature_store(w), batch_size=8, splits=["train"], tokenizer=tok)
    H.warm_candidate_buffer(renv)
    spk = ref_speaker.Speaker(cfg, torch.device("cpu"), tok, env=OldKeys(renv))
    mods = [spk.encoder, spk.decoder]
    out = {"world": dict(n_scans=3, seed=1, n_items=40, B=8), "max_decode": 24,
           "w_checksum": [float(p.detach().double().sum()) for m in mods for p in m.parameters()], "batches": []}
    random.seed(1)
    for _ in range(2):
        obs = spk.env.reset()
        (img, can), lens = spk.from_shortest_path()
        insts = torch.from_numpy(np.array([ob["instr_encoding"] for ob in obs]))
        spk.env.reset(restart=True)
        loss, wa, sa = spk.teacher_forcing(train=False)
        scores = spk.teacher_forcing(train=False, features=((img, can), lens), insts=insts, for_listener=True).detach()
        for m in mods:
            m.zero_grad()
        spk.encoder.eval(), spk.decoder.eval()
        # gradients of the eval-mode loss (train=True would switch torch's dropout on): the class's own modules and criterion
        ctx = spk.encoder(can.clone(), img.clone(), lens, already_dropfeat=True)
        z = torch.zeros(1, len(lens), 512)
        logits, _, _ = spk.decoder(insts, ctx, ref_speaker.utils.length2mask(lens, torch.device("cpu")), z, z)
        l2 = spk.softmax_loss(input=logits.permute(0, 2, 1)[:, :, :-1], target=insts[:, 1:])
        l2.backward()
        gn = [float(p.grad.norm()) if p.grad is not None else 0.0 for m in mods for p in m.parameters()]
        spk.env.reset(restart=True)
        words = spk.infer_batch()
        out["batches"].append(dict(instr_ids=[ob["instr_id"] for ob in obs], lengths=np.asarray(lens).tolist(), loss=float(loss),
                                   word_accu=float(wa), sent_accu=float(sa), scores=scores, loss_grad=float(l2),
                                   grad_norms=gn, words=torch.from_numpy(words.astype(np.int64)),
                                   can_sum=float(can.double().sum()), img_sum=float(img.double().sum())))
    return out


def beam_cases():
    """The real EnvDrop, Follower and Self-Monitor agents' ``_dijkstra`` (base.py:183-397; K = 3) on the seeded synthetic world, eval mode:
    per episode the K best paths (poses, actions, listener scores) and the navigation path; world and weights are regenerated
    from the seeds by the test."""
    import clvln_b200  # noqa: F401
    from clvln_b200.environ import make_world, make_items
    from oracle import ref_harness as H
    w = make_world(n_scans=3, seed=1)
    items = make_items(w, 40, seed=1)
    H.install(w, {"train": items})
    import src.agent as agent_mod
    import src.environ as environ
    tok = H.StubTokenizer(items)
    fs = H.feature_store(w)
    dev = torch.device("cpu")
    out = {"world": dict(n_scans=3, seed=1, n_items=40, B=6), "K": 3}
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
        mods = [ag.encoder, ag.decoder] + ([ag.critic] if kind == "ENVDROP" else [])
        with torch.no_grad():
            res = ag._dijkstra(3)
        out[kind] = dict(w_checksum=[float(p.detach().double().sum()) for m in mods for p in m.parameters()],
                         results=[dict(instr_id=r["instr_id"], dijk_path=list(r["dijk_path"]),
                                       paths=[dict(trajectory=[tuple(x) for x in p["trajectory"]], action=list(p["action"]),
                                                   listener_scores=[float(x) for x in p["listener_scores"]],
                                                   listener_actions=list(p["listener_actions"])) for p in r["paths"]])
                                  for r in res])
    return out


def eval_cases():
    """Random-walk trajectories on a seeded synthetic world scored by the reference's own Evaluation.score
    (src/engine/evaluator.py:101-146): the summary and the per-trajectory lists travel as tests/golden/eval.json."""
    import clvln_b200  # noqa: F401
    from clvln_b200.environ import make_items, make_world
    from oracle import ref_harness as H
    w = make_world(n_scans=3, seed=4)
    items = make_items(w, 60, seed=4, instr_per_path=3)
    H.install(w, {"val": items})
    import src.engine.evaluator as RE
    ref_eval = RE.Evaluation(["val"], data_name="R2R")
    rng = random.Random(7)
    results = []
    for it in items:
        traj = [it["path_g"][0]]
        if rng.random() < 0.5:
            traj = list(it["path_g"][:rng.randint(1, len(it["path_g"]))])
        for _ in range(rng.randint(0, 6)):
            g = traj[-1]
            traj.append(int(w.cand_vp[g, rng.randrange(int(w.n_cand[g]))]))
        s = it["scan_idx"]
        o = int(w.scan_off[s])
        results.append({"instr_id": it["instr_id"], "trajectory": [[w.vp_names[s][g - o], 0.0, 0.0] for g in traj]})
    summary, scores = ref_eval.score(results)
    return {"world": {"n_scans": 3, "seed": 4}, "items": {"n": 60, "seed": 4, "instr_per_path": 3}, "results": results,
            "summary": {k: float(v) for k, v in summary.items()},
            "scores": {k: [float(x) for x in v] for k, v in scores.items()}}


def main():
    from oracle import ref_loader
    assert ref_loader.reference_available(), "needs /root/reference"
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    if only in ("speaker", "beam"):                             # only that fixture (the others are unchanged)
        torch.save({"speaker": speaker_cases, "beam": beam_cases}[only](), os.path.join(OUT, only + ".pt"))
        print(only + ".pt", os.path.getsize(os.path.join(OUT, only + ".pt")))
        return
    torch.save(speaker_cases(), os.path.join(OUT, "speaker.pt"))
    torch.save(beam_cases(), os.path.join(OUT, "beam.pt"))
    torch.save(module_cases(), os.path.join(OUT, "modules.pt"))
    torch.save(rollout_cases(), os.path.join(OUT, "rollouts.pt"))
    import json
    with open(os.path.join(OUT, "eval.json"), "w") as f:
        json.dump(eval_cases(), f)
    for f in os.listdir(OUT):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
