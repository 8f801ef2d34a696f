"""Rollout-level parity on the GPU: the product agents (device-resident rollouts through the
C-ABI kernels) against the oracle's restatement of the reference agents (oracle/port_rollout.py,
itself pinned to the real reference by tests/test_oracle_vs_reference.py and tests/golden).

Bars (north star): logits and loss max-relative error <= 1e-3, gradient cosine >= 0.9999,
teacher-forced argmax agreement >= 99.9 %; trajectories (integer state) bit-exact.
Train-mode runs feed the oracle the very keep-masks the kernels drew (Philox streams are
regenerated from the recorded call sites), and sampled actions are replayed as forced actions.
"""
import copy
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False

KINDS = {"ENVDROP": 12, "FOLLOWER": 10, "SELF-MONITOR": 10}


def _setup(kind, B=8, n_items=40, seed=1, fixed_len=None):
    import clvln_b200
    from clvln_b200 import utils
    from clvln_b200.agent import build_agent
    from clvln_b200.environ import make_world, make_items, R2RBatch
    from oracle import port_env as PE, port_rollout as PR
    dev = torch.device("cuda:0")
    world = make_world(n_scans=3, seed=seed)
    items = make_items(world, n_items, seed=seed, fixed_len=fixed_len)
    cfg = utils.agent_cfg(kind)
    cfg.AGENT.MAX_EPISODE_LEN = KINDS[kind]
    random.seed(2020)
    env = R2RBatch(world, items, batch_size=B, device=dev)
    torch.manual_seed(2020)
    agent = build_agent(cfg, utils.StubTokenizer(), dev)            # re-seeds random to 1 (base.py:28)
    agent.env = env
    agent.sync_every = 1
    random.seed(2020)
    penv = PE.R2RBatchPort(PE.WorldView(world), items, batch_size=B)
    random.seed(1)
    assert [d["instr_id"] for d in penv.data] == [d["instr_id"] for d in env.data]
    mods = agent._modules()
    sds = [{k: v.detach().cpu().clone().requires_grad_(v.dtype.is_floating_point and "running" not in k
                                                       and k != "position.pe")
            for k, v in m.state_dict().items()} for m in mods]
    mc = {"ENVDROP": cfg.MODEL.ENVDROP, "FOLLOWER": cfg.MODEL.FOLLOWER, "SELF-MONITOR": cfg.MODEL.MONITOR}[kind]
    pag = PR.Agent(kind, sds[0], sds[1], sds[2] if len(sds) > 2 else None, hidden=mc.HIDDEN_SIZE,
                   bidirectional=mc.ENC_BIDIRECTION, enc_layers=mc.ENC_LAYERS, episode_len=KINDS[kind])
    return agent, pag, env, penv, sds, cfg


def _grads(params):
    return torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).detach().flatten().cpu().double()
                      for p in params])


def _cos(a, b):
    return float(torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-30))


def _rel(a, b):
    a, b = torch.as_tensor(a).detach().cpu().double(), torch.as_tensor(b).detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))


def _compare_traces(mine, theirs):
    """Per-step logits (finite slots), -inf pattern, targets; returns teacher-argmax agreement."""
    agree = tot = 0
    for m, o in zip(mine, theirs):
        C = o["logits"].shape[1]
        lm = m["logits"][:, :C].cpu()
        assert torch.equal(torch.isinf(lm), torch.isinf(o["logits"]))
        if lm.shape[1] < m["logits"].shape[1]:
            assert bool(torch.isinf(m["logits"][:, C:]).all())
        assert torch.equal(m["target"].cpu().long(), o["target"].long())
        live = o["target"] >= 0
        fin = ~torch.isinf(o["logits"])
        if live.any():
            sel = fin & live.unsqueeze(1)
            assert _rel(lm[sel], o["logits"][sel]) < 1e-3
            agree += int((lm[live].argmax(1) == o["logits"][live].argmax(1)).sum())
            tot += int(live.sum())
    return agree, tot


def _mask_feed(agent, cand_widths=None):
    """Regenerate the kernels' keep-masks from the recorded call sites, grouped by oracle tag."""
    from clvln_b200 import ops
    feed = {}
    for tag, shape, p, off in agent.rng.log:
        feed.setdefault(tag, []).append(ops.dropout_mask(shape, p, agent.rng, off).cpu())
    return feed


@pytest.mark.parametrize("fused", [True, False], ids=["fused", "modules"])
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_envdrop_rollouts_match_oracle(mode, fused):
    """fused: the agent's hand-differentiated decoder rollout (agent/fused.py);
    modules: the same rollout through the drop-in nn.Modules (one autograd node per op)."""
    from oracle import port_modules as P, port_rollout as PR
    agent, pag, env, penv, sds, cfg = _setup("ENVDROP")
    agent.fused = fused
    getattr(agent, mode)()
    agent.rng.log = [] if mode == "train" else None
    agent.rng.begin_iteration()
    # ---- product: teacher rollout (IL) + sampled rollout (A2C) on the same batch, one backward
    agent.trace = []
    agent.rollout(train_ml=True, train_rl=False, feedback="teacher")
    ml, tr1, st1 = agent.loss["ml_loss"], agent.trace, agent.last_state
    agent.trace = []
    agent.rollout(train_ml=False, train_rl=True, restart=True, feedback="sample")
    rl, tr2, st2 = agent.loss["rl_loss"], agent.trace, agent.last_state
    (ml + rl).backward()
    g_mine = _grads(agent.trainable_params())
    # ---- oracle: same batch, same masks, the sampled actions replayed
    drop = None
    if mode == "train":
        feed = _mask_feed(agent)
        # critic masks: the oracle calls critic(last_h) first, then hidden_states[t] for t = T-1..0
        last, vals = feed["critic"][0], feed["critic"][1]
        n = len(tr2)
        vals = vals.view(n, -1, vals.shape[-1])
        feed["critic"] = [last] + [vals[t] for t in range(n - 1, -1, -1)]
        steps = iter(tr1 + tr2 + [None])
        # candidate masks are drawn for 16 slots; the oracle's tensor is max(n_cand)+1 wide per step
        feed["cand"] = [m for m in feed["cand"]]
        drop = _WidthDrop(feed)
    forced = [t["action"].cpu().numpy() for t in tr2]
    _, l1 = PR.rollout_envdrop(pag, penv, train_ml=True, train_rl=False, feedback="teacher", drop=drop)
    otr1 = pag.trace["steps"]
    _, l2 = PR.rollout_envdrop(pag, penv, train_ml=False, train_rl=True, restart=True, feedback=forced, drop=drop)
    otr2 = pag.trace["steps"]
    (l1["ml_loss"] + l2["rl_loss"]).backward()
    g_ref = _grads([v for sd in sds for v in sd.values() if v.requires_grad])
    # ---- compare
    assert len(tr1) == len(otr1) and len(tr2) >= len(otr2)
    a1, n1 = _compare_traces(tr1, otr1)
    _compare_traces(tr2[:len(otr2)], otr2)
    assert a1 / n1 >= 0.999
    assert _rel(ml, l1["ml_loss"]) < 1e-3 and _rel(rl, l2["rl_loss"]) < 1e-3
    assert g_mine.shape == g_ref.shape
    assert _cos(g_mine, g_ref) >= 0.9999
    assert _rel(g_mine, g_ref) < 5e-3


class _WidthDrop:
    """oracle Drop that crops an injected mask to the tensor it is applied to (the kernels draw
    candidate masks for all 16 slots; the oracle's candidate tensor is only max(n_cand)+1 wide)."""

    def __init__(self, masks):
        self.mode = "masks"
        self.masks = {k: list(v) for k, v in masks.items()}

    def __call__(self, x, p, tag):
        if p <= 0.0:
            return x
        keep = self.masks[tag].pop(0)
        if keep.dim() == 3 and x.dim() == 2:            # [B, 16 slots, n] mask for the oracle's flattened [B * C, n] rows
            keep = keep[:, :x.shape[0] // keep.shape[0]].reshape(x.shape[0], -1)
        if keep.shape != x.shape:
            keep = keep[tuple(slice(0, s) for s in x.shape)]
        return x * keep.to(x.dtype) * (1.0 / (1.0 - p))


@pytest.mark.parametrize("kind", ["FOLLOWER", "SELF-MONITOR"])
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_follower_monitor_rollouts_match_oracle(kind, mode):
    from oracle import port_rollout as PR
    agent, pag, env, penv, sds, cfg = _setup(kind)
    getattr(agent, mode)()
    pag.training = mode == "train"
    agent.rng.log = [] if mode == "train" else None
    agent.rng.begin_iteration()
    agent.trace = []
    agent.rollout(feedback="teacher")
    l_t, tr1 = agent.ml_loss, agent.trace
    agent.trace = []
    agent.rollout(feedback="sample", train_cl=True)
    l_s, tr2 = agent.ml_loss, agent.trace
    (l_t + l_s.sum()).backward()
    g_mine = _grads(agent.trainable_params())
    drop = None
    if mode == "train":
        feed = _mask_feed(agent)
        if "mlp" in feed:           # MLPwithBN runs twice per step: previous action, then the candidates
            feed["mlp_prev"], feed["mlp_cand"] = feed["mlp"][0::2], feed["mlp"][1::2]
            # the product runs all 16 candidate slots through the MLP (device-side width, masked BatchNorm statistics);
            # the oracle's tensor is [B * max(n_cand + 1), 1024]: hand the masks over per (episode, slot)
            nb = feed["mlp_prev"][0].shape[0]
            feed["mlp_cand"] = [m.view(nb, -1, m.shape[-1]) for m in feed["mlp_cand"]]
        drop = _WidthDrop(feed)
    forced = [t["action"].cpu().numpy() for t in tr2]
    roll = PR.rollout_follower if kind == "FOLLOWER" else PR.rollout_monitor
    o1 = roll(pag, penv, feedback="teacher", drop=drop)
    otr1 = pag.trace["steps"]
    o2 = roll(pag, penv, feedback=forced, train_cl=True, drop=drop)
    otr2 = pag.trace["steps"]
    (o1[1] + o2[1].sum()).backward()
    g_ref = _grads([v for sd in sds for v in sd.values() if v.requires_grad])
    assert len(tr1) == len(otr1) and len(tr2) >= len(otr2)
    a1, n1 = _compare_traces(tr1, otr1)
    _compare_traces(tr2[:len(otr2)], otr2)
    assert a1 / n1 >= 0.999
    assert _rel(l_t, o1[1]) < 1e-3 and _rel(l_s, o2[1]) < 1e-3
    assert _cos(g_mine, g_ref) >= 0.9999
    if kind == "SELF-MONITOR" and mode == "train":                  # BatchNorm running statistics
        dsd = agent.decoder.state_dict()
        for k in ("proj_navigable_mlp.mlp.0.running_mean", "proj_navigable_mlp.mlp.2.running_var"):
            assert _rel(dsd[k], sds[1][k]) < 1e-4


def test_trajectories_match_obs_dict_face():
    """Device rollouts (argmax) produce the trajectories the obs-dict face + oracle agent produce."""
    from oracle import port_rollout as PR
    agent, pag, env, penv, sds, cfg = _setup("ENVDROP")
    agent.eval()
    traj = agent.rollout(train_ml=False, feedback="argmax")
    otraj, _ = PR.rollout_envdrop(pag, penv, train_ml=False, feedback="argmax")
    assert [t["instr_id"] for t in traj] == [t["instr_id"] for t in otraj]
    for a, b in zip(traj, otraj):
        assert [p[0] for p in a["path"]] == [p[0] for p in b["path"]]
        assert np.allclose([p[1:] for p in a["path"]], [p[1:] for p in b["path"]])


def test_train_step_and_flat_optimizer():
    """One EnvDrop TrainStep == reference recipe (clip encoder/decoder to 40, RMSprop) applied to the
    same gradients; parameters stay views of the flat buffer; the loss goes down over a few steps."""
    from clvln_b200.engine import TrainStep
    agent, pag, env, penv, sds, cfg = _setup("ENVDROP", B=8)
    agent.train()
    step = TrainStep(cfg, agent)
    ref = [p.detach().clone().requires_grad_(True) for p in agent.trainable_params()]
    opt = torch.optim.RMSprop(ref, lr=cfg.TRAIN.LR)
    n_enc = len(list(agent.encoder.parameters()))
    n_dec = len(list(agent.decoder.parameters()))
    losses = []
    for it in range(3):
        before = [p.detach().clone() for p in agent.trainable_params()]
        losses.append(float(step()))
        for r, p, b in zip(ref, agent.trainable_params(), before):
            assert torch.equal(r.detach(), b) or _rel(r.detach(), b) < 1e-5
            r.data.copy_(b)
            r.grad = p.grad.detach().clone()
        torch.nn.utils.clip_grad_norm_(ref[:n_enc], 40.0)
        torch.nn.utils.clip_grad_norm_(ref[n_enc:n_enc + n_dec], 40.0)
        opt.step()
        for r, p in zip(ref, agent.trainable_params()):
            assert _rel(p, r) < 1e-4
    assert all(np.isfinite(losses))
    flat = step.opt.flat
    for p in agent.trainable_params():
        assert flat.data_ptr() <= p.data_ptr() < flat.data_ptr() + flat.numel() * 4


@pytest.mark.parametrize("train_cl", [False, True], ids=["sum", "per_item"])
def test_paired_rollouts_equal_separate_rollouts(train_cl):
    """agent.rollout_pair (teacher-forced + sampled rollout of one iteration stepped as one batch of 2B episodes,
    the teacher half dropping out after its last step) == the two rollout() calls of trainer.py:411-421: same
    imitation / A2C losses and gradients.  Dropout off so both paths are the same function of the weights (each
    half draws its own masks otherwise); the sampled actions use the same Philox offsets in both paths."""
    out = []
    for pair in (False, True):
        agent, pag, env, penv, sds, cfg = _setup("ENVDROP", B=8)
        agent.train()
        agent.sync_every = 0
        agent.encoder.drop_ratio = agent.decoder.drop_ratio = agent.decoder.feat_drop_ratio = 0.0
        agent.critic.state2value[2].p = 0.0
        if pair:
            agent.rollout_pair(train_cl=train_cl)
        else:
            agent.rollout(train_ml=True, train_rl=False, train_cl=train_cl, feedback="teacher")
            ml = agent.loss["ml_loss"]
            agent.rollout(train_ml=False, train_rl=True, train_cl=train_cl, restart=True, feedback="sample")
            agent.loss = {"ml_loss": ml, "rl_loss": agent.loss["rl_loss"]}
        ml, rl = agent.loss["ml_loss"], agent.loss["rl_loss"]
        (ml + rl).sum().backward()
        out.append((ml.detach().cpu(), rl.detach().cpu(), _grads(agent.trainable_params()),
                    agent.last_state.vp[:, :8].cpu(), [float(x) for x in agent.logs["total"]]))
    (ml0, rl0, g0, vp0, tot0), (ml1, rl1, g1, vp1, tot1) = out
    assert torch.equal(vp0, vp1[:vp0.shape[0]])                      # the sampled half walked the same trajectories
    assert tot0 == tot1
    # the A2C loss is a small difference of larger terms (policy, critic, entropy): absolute floor for the split-K
    # summation-order noise of the GEMMs
    assert _rel(ml1, ml0) < 1e-4 and torch.allclose(rl1, rl0, rtol=1e-3, atol=2e-5)
    assert _cos(g1, g0) > 0.99999 and _rel(g1, g0) < 1e-3


def test_graph_replay_tracks_the_optimiser():
    """CUDA-graph replays of the iteration must see the weights the fused optimiser wrote (the bf16
    weight splits are re-derived inside the graph): graphed steps == eager steps.  Dropout off and
    teacher forcing only, so both runs are deterministic functions of the weights."""
    from clvln_b200.engine import TrainStep
    from clvln_b200.engine.graphs import GraphedTrainStep
    finals = []
    for graphed in (False, True):
        agent, pag, env, penv, sds, cfg = _setup("ENVDROP", B=8, fixed_len=20)
        agent.train()
        agent.encoder.drop_ratio = agent.decoder.drop_ratio = agent.decoder.feat_drop_ratio = 0.0
        cfg.AGENT.FEEDBACK = "teacher"
        cfg.TRAIN.LR = 3e-4
        init = [p.detach().clone() for p in agent.trainable_params()]
        step = (GraphedTrainStep if graphed else TrainStep)(cfg, agent)
        losses = [float(step()) for _ in range(4)]
        finals.append((losses, init, [p.detach().clone() for p in agent.trainable_params()]))
    (l0, i0, p0), (l1, _, p1) = finals
    assert np.allclose(l0, l1, rtol=1e-2, atol=1e-4), (l0, l1)     # round-off is amplified by RMSprop's g/sqrt(v)
    assert max(_rel(a, b) for a, b in zip(p0, i0)) > 1e-3           # the optimiser did move the weights
    # RMSprop turns round-off in near-zero gradients into +-lr steps, so compare against the distance moved
    num = sum(float((a - b).norm()) ** 2 for a, b in zip(p0, p1)) ** 0.5
    den = sum(float((a - i).norm()) ** 2 for a, i in zip(p0, i0)) ** 0.5
    assert num <= 0.25 * den, (num, den)


def _cl_setup(kind, B, n_items=120):
    import clvln_b200  # noqa: F401
    from clvln_b200 import utils
    from clvln_b200.agent import build_agent
    from clvln_b200.environ import make_world, make_items, split_rounds
    dev = torch.device("cuda:0")
    world = make_world(n_scans=3, seed=2)
    items = make_items(world, n_items, seed=2)
    cfg = utils.agent_cfg(kind)
    cfg.AGENT.MAX_EPISODE_LEN = KINDS[kind]
    cfg.TRAIN.BATCH_SIZE = B
    cfg.TRAIN.MAX_EPOCH, cfg.TRAIN.ITER_PER_EPOCH = 3, 3
    torch.manual_seed(2020)
    agent = build_agent(cfg, utils.StubTokenizer(), dev)
    return world, split_rounds(items), cfg, agent, dev


def test_self_paced_envdrop_trainer_end_to_end():
    """BASELINE config 4 in small: EnvDrop + TRAIN.CLMODE=SELF-PACE through build_trainer(...).train (graph-replayed
    iterations): per-item losses land at the visited dataset indices, the pace update of curriculum.py:402-448 runs on
    the device-resident weight vector the captured graph reads, parameters move, losses stay finite."""
    from clvln_b200.engine import build_trainer
    from clvln_b200.engine.graphs import GraphedTrainStep
    from clvln_b200.environ import CLR2RBatch
    world, rounds, cfg, agent, dev = _cl_setup("ENVDROP", B=8)
    cfg.TRAIN.CLMODE = "SELF-PACE"
    sp = cfg.TRAIN.SELF_PACE
    sp.FUNC, sp.LAMB, sp.MIU, sp.WCTRL, sp.CRATE, sp.INTERVAL, sp.BURN_IN, sp.STRATEGY = "linear", 2.0, 2.0, 0.5, 1.0, 1, 1, "epoch"
    random.seed(2020)
    env = CLR2RBatch(world, rounds, batch_size=8, c_rate=sp.CRATE, device=dev)
    trainer = build_trainer(cfg, env, dev)
    assert isinstance(trainer.make_step(cfg, agent), GraphedTrainStep)
    w0 = trainer.weight.clone()
    wptr = trainer.weight.data_ptr()
    p0 = [p.detach().clone() for p in agent.trainable_params()]
    visited = set()
    orig = trainer.record

    def record(index, item):
        visited.update(index.tolist())
        assert item.shape == index.shape and bool(torch.isfinite(item).all())
        orig(index, item)
    trainer.record = record
    trainer.train(cfg, agent, None, env, None, log=lambda *_: None)
    assert len(trainer.history) == 3 and all(np.isfinite(h["loss_sum"]) for h in trainer.history)
    nz = set(torch.nonzero(trainer.loss_for_item).flatten().tolist())
    assert nz and nz <= visited and len(visited) <= 3 * 3 * 8
    assert trainer.weight.data_ptr() == wptr and not torch.equal(trainer.weight, w0)      # updated in place
    assert float(trainer.weight.min()) > 0.0
    assert max(_rel(a, b) for a, b in zip(agent.trainable_params(), p0)) > 1e-4


def test_naive_curriculum_monitor_trainer_end_to_end():
    """BASELINE config 3 in small: Self-Monitor + TRAIN.CLMODE=NAIVE (round switch every epoch here) through
    build_trainer(...).train on the drop-in module path."""
    from clvln_b200.engine import build_trainer, NaiveCurriculum
    from clvln_b200.environ import R2RBatch
    world, rounds, cfg, agent, dev = _cl_setup("SELF-MONITOR", B=8)
    cfg.TRAIN.CLMODE = "NAIVE"
    random.seed(2020)
    envs = {f"round_{k}": R2RBatch(world, [it for j in range(1, k + 1) for it in rounds[j]], batch_size=8, device=dev)
            for k in range(1, 6)}
    trainer = build_trainer(cfg, envs, dev)
    assert isinstance(trainer, NaiveCurriculum)
    trainer.switch_epoch = 1
    seen = []
    pick = trainer.pick_env
    trainer.pick_env = lambda te, ep: seen.append(pick(te, ep)) or seen[-1]
    p0 = [p.detach().clone() for p in agent.trainable_params()]
    trainer.train(cfg, agent, None, envs, None, log=lambda *_: None)
    assert seen == [envs["round_1"], envs["round_2"], envs["round_3"]]
    assert all(np.isfinite(h["loss_sum"]) for h in trainer.history)
    assert max(_rel(a, b) for a, b in zip(agent.trainable_params(), p0)) > 1e-4


def test_bench_cuda_arm_prints_the_contract_line():
    """`bench.py` (small world, 2 steps): one JSON line with the contract's keys, device-side launches counted,
    an end-to-end block with per-step H2D/D2H bytes, the clocks sample and the panorama-attention roofline."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--small", "--steps", "2", "--warmup", "3",
                          "--no-cpu-baseline"], capture_output=True, text=True, timeout=900, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] >= 3 and d["value"] > 0
    assert d["gpu_launches"] > 100 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3


def test_eval_loop_results_and_checkpoint_round_trip(tmp_path):
    """The rest of the agent protocol main.py / the trainers use (base.py:63-112, envdrop.py:298-313): `test()` rolls the
    validation env out with argmax actions until an instruction repeats, `get_results()` / `write_results()` hold one
    trajectory per instr_id (starting at the episode's start viewpoint), the evaluator scores them, and
    save_model -> load_model into a fresh agent reproduces the same trajectories."""
    import json
    from clvln_b200 import utils
    from clvln_b200.agent import build_agent
    from clvln_b200.engine import evaluate
    agent, pag, env, penv, sds, cfg = _setup("ENVDROP", B=8, n_items=20)
    agent.results_save_dir = str(tmp_path)
    agent.eval()
    agent.test(iters=None, feedback="argmax")
    res = agent.get_results()
    ids = [r["instr_id"] for r in res]
    assert sorted(ids) == sorted(it["instr_id"] for it in env.data) and len(set(ids)) == len(ids)
    start = {it["instr_id"]: it["path"][0] for it in env.data}
    assert all(r["trajectory"][0][0] == start[r["instr_id"]] for r in res)
    scores = evaluate(env, res)
    for k in ("nav_error", "oracle_error", "steps", "lengths", "spl", "ndtw", "sdtw", "cls", "success_rate", "oracle_rate"):
        assert np.isfinite(scores[k]), k
    agent.write_results("val")
    assert len(json.load(open(tmp_path / "val.json"))) == len(res)
    ckpt = str(tmp_path / "ckpt.pt")
    agent.save_model(ckpt, cfg=cfg, last_epoch=3)
    torch.manual_seed(7)                                             # different init: the checkpoint must overwrite it
    other = build_agent(cfg, utils.StubTokenizer(), agent.device)
    meta = other.load_model(ckpt)
    assert meta["last_epoch"] == 3 and set(meta) >= {"encoder_state_dict", "decoder_state_dict", "critic_state_dict"}
    other.env = env
    other.eval()
    other.test(iters=None, feedback="argmax")
    again = {r["instr_id"]: r["trajectory"] for r in other.get_results()}
    assert all([tuple(p) for p in again[r["instr_id"]]] == [tuple(p) for p in r["trajectory"]] for r in res)


def test_classic_trainer_with_validation_and_checkpoints(tmp_path):
    """ClassicTrainer.train with a validation env and the evaluator (trainer.py:481-511): every EVAL_INTERVAL epochs the
    agent is tested with argmax actions, scored, and the best / latest checkpoints are written the way the reference names
    them; a run resumed from `latest` starts at last_epoch + 1 (trainer.py:378)."""
    import os
    from clvln_b200.engine import build_trainer, evaluate
    from clvln_b200.environ import R2RBatch
    world, rounds, cfg, agent, dev = _cl_setup("ENVDROP", B=8, n_items=60)
    items = [it for k in range(1, 6) for it in rounds[k]]
    random.seed(2020)
    train_env = R2RBatch(world, items[:40], batch_size=8, device=dev)
    val_env = R2RBatch(world, items[40:], batch_size=8, name="val_seen", device=dev)
    cfg.TRAIN.MAX_EPOCH, cfg.TRAIN.ITER_PER_EPOCH, cfg.TRAIN.EVAL_INTERVAL = 2, 2, 1
    cfg.OUTPUT.CKPT_DIR = str(tmp_path)
    trainer = build_trainer(cfg, train_env, dev)
    trainer.train(cfg, agent, None, train_env, {"val_seen": val_env}, evaluator=evaluate, log=lambda *_: None)
    assert len(trainer.history) == 2 and all("val_seen" in h and 0.0 <= h["val_seen"]["success_rate"] <= 1.0 for h in trainer.history)
    files = os.listdir(tmp_path)
    latest = [f for f in files if f.startswith("latest_avgloss:")]
    assert len(latest) == 1
    assert all(np.isfinite(h["loss_avg"]) for h in trainer.history)
    ckpt = torch.load(os.path.join(tmp_path, latest[0]), map_location="cpu", weights_only=False)
    assert ckpt["last_epoch"] == 2 and "critic_state_dict" in ckpt
    # resume: two more epochs start at epoch 3
    cfg.OUTPUT.RESUME = latest[0][:-3]
    cfg.TRAIN.MAX_EPOCH = 4
    trainer2 = build_trainer(cfg, train_env, dev)
    trainer2.train(cfg, agent, None, train_env, None, log=lambda *_: None)
    assert [h["epoch"] for h in trainer2.history] == [3, 4]


def test_device_evaluation_matches_host_scores():
    """engine.Evaluation.score_device (ONE launch of vln_eval_paths over the whole split: nav / oracle error, SPL, nDTW,
    SDTW, CLS per trajectory in float64 on the HBM distance table) == Evaluation.score (the host restatement that
    tests/_ref_check_eval.py pins against the reference's Evaluation.score and whose DTW / CLS reproduce the reference's
    doctest values) on random walks of every length, incl. trajectories that never leave the start."""
    from clvln_b200.engine import Evaluation
    agent, pag, env, penv, sds, cfg = _setup("ENVDROP", B=8, n_items=60)
    w = env.world
    rs = np.random.RandomState(3)
    results = []
    for it in env.data:
        g = it["path_g"][0]
        traj = [(env._vp_name(g), 0.0, 0.0)]
        for _ in range(int(rs.randint(0, 12))):
            n = int(w.n_cand[g])
            if n == 0:
                break
            g = int(w.cand_vp[g, rs.randint(0, n)])
            traj.append((env._vp_name(g), 0.0, 0.0))
        if rs.rand() < 0.3:                                   # some follow the ground truth exactly
            traj = [(env._vp_name(x), 0.0, 0.0) for x in it["path_g"]]
        results.append({"instr_id": it["instr_id"], "trajectory": traj})
    ev = Evaluation(env)
    host, per = ev.score(results)
    devs, m = ev.score_device(results, agent.store_of(env))
    for k in host:
        assert abs(host[k] - devs[k]) <= 1e-9 * max(1.0, abs(host[k])), (k, host[k], devs[k])
    assert np.allclose(m[:, 5], per["ndtws"], rtol=1e-12, atol=0) and np.allclose(m[:, 7], per["clss"], rtol=1e-12, atol=0)
