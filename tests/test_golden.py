"""Pin the oracle (oracle/port_*.py) against golden vectors produced by the REAL reference
(oracle/make_golden.py, run in the build container; fixtures under tests/golden/).  Runs anywhere
(CPU) — this is what carries the pin to machines without /root/reference."""
import os
import random

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
G = os.path.join(HERE, "golden")


def close(a, b, tol=1e-5):
    d = (a - b).abs().max().item()
    assert d <= tol * max(1.0, b.abs().max().item()), d


@pytest.fixture(scope="module")
def mods():
    return torch.load(os.path.join(G, "modules.pt"), weights_only=False)


def test_encoder_variants(mods):
    from oracle import port_modules as P
    for name in ("enc_bi1", "enc_bi2", "enc_uni"):
        c = mods[name]
        E, H, bi, nl = c["cfg"]
        ctx, h, cc = P.encoder_lstm(c["sd"], c["toks"], c["lens"], bidirectional=bi, num_layers=nl, drop_ratio=0.5)
        close(ctx, c["ctx"]), close(h, c["h"]), close(cc, c["c"])


def test_decoders_and_critic(mods):
    from oracle import port_modules as P
    c = mods["envdrop"]
    lo, (h1, c1), ht, _ = P.envdrop_decoder(c["sd"], c["a"], c["img"], c["cand"], c["ht"], c["c0"], c["ctx"], c["mask"])
    close(lo, c["logit"], 1e-4), close(h1, c["h1"]), close(c1, c["c1"]), close(ht, c["h_tilde"])
    c = mods["follower"]
    lo, (h1, c1), (ac, av) = P.follower_decoder(c["sd"], c["img"], c["ap"], c["cand"], c["h0"], c["c0"], c["ctx"], c["mask"])
    close(lo, c["logit"], 1e-4), close(h1, c["h1"]), close(ac, c["alpha_c"]), close(av, c["alpha_v"])
    for training in (0, 1):
        c = mods[f"monitor_train{training}"]
        sd = {k: v.clone() for k, v in c["sd"].items()}
        (lo, pr), (h1, c1), (ca, va) = P.monitor_decoder(sd, c["ap"], c["cand"], c["h0"], c["c0"], c["ctx"], c["mask"],
                                                         c["cmask"], training=bool(training))
        close(lo, c["logit"], 2e-4), close(pr, c["prog"], 1e-4), close(h1, c["h1"], 1e-4), close(va, c["cand_attn"], 1e-4)
        close(sd["proj_navigable_mlp.mlp.0.running_mean"], c["rm"]), close(sd["proj_navigable_mlp.mlp.2.running_var"], c["rv"], 1e-4)
    c = mods["critic"]
    close(P.critic(c["sd"], c["x"]), c["y"])


@pytest.fixture(scope="module")
def roll():
    return torch.load(os.path.join(G, "rollouts.pt"), weights_only=False)


def _world(roll):
    import clvln_b200  # noqa: F401
    from clvln_b200.environ import make_world, make_items
    w = roll["world"]
    world = make_world(n_scans=w["n_scans"], seed=w["seed"])
    return world, make_items(world, w["n_items"], seed=w["seed"])


def test_minibatch_order_matches_reference(roll):
    """Product env and oracle env draw the reference's minibatches (incl. wrap-around reshuffles)."""
    from clvln_b200.environ import R2RBatch
    from oracle import port_env as PE
    world, items = _world(roll)
    for make in (lambda: R2RBatch(world, items, batch_size=16),
                 lambda: PE.R2RBatchPort(PE.WorldView(world), items, batch_size=16)):
        random.seed(2020)
        env = make()
        random.seed(1)
        got = []
        for _ in range(7):
            env._next_minibatch()
            got.append([it["instr_id"] for it in env.batch])
        assert got == roll["order"]


@pytest.mark.parametrize("kind", ["ENVDROP", "FOLLOWER", "MONITOR"])
def test_port_rollouts_match_reference_golden(roll, kind):
    """Oracle rollouts on regenerated world + weights reproduce the real agents' losses,
    gradient norms and trajectories (eval mode; sampling replays torch's RNG stream)."""
    from clvln_b200 import utils
    from clvln_b200.model import EncoderLSTM, EnvDropDecoder, AttnDecoderLSTM, MonitorDecoder, Critic
    from oracle import port_env as PE, port_rollout as PR
    world, items = _world(roll)
    g = roll[kind]
    random.seed(2020)
    torch.manual_seed(2020)
    penv = PE.R2RBatchPort(PE.WorldView(world), items, batch_size=8)
    # the reference builds encoder, decoder(, critic) in this order under the same seed; the product
    # modules have the same parameter containers and init order, so the weights come out identical
    if kind == "ENVDROP":
        mods = [EncoderLSTM(992, 256, 512, 0, 0.5, True, 1), EnvDropDecoder(512, 0.5, 0.3, 64, 128, 2176), Critic(512, 0.5)]
        kw = dict(hidden=512, bidirectional=True, enc_layers=1, episode_len=12)
    elif kind == "FOLLOWER":
        mods = [EncoderLSTM(992, 300, 256, 0, 0.5, True, 2), AttnDecoderLSTM(256, 0.5, 2176, 2176)]
        kw = dict(hidden=256, bidirectional=True, enc_layers=2, episode_len=10)
    else:
        mods = [EncoderLSTM(992, 256, 512, 0, 0.5, False, 1), MonitorDecoder(512, 0.5, 80, [1024], 2176, 2176)]
        kw = dict(hidden=512, bidirectional=False, enc_layers=1, episode_len=10)
    chk = [float(p.detach().double().sum()) for m in mods for p in m.parameters()]
    assert len(chk) == len(g["w_checksum"]) and max(abs(a - b) for a, b in zip(chk, g["w_checksum"])) < 1e-6, \
        "regenerated weights differ from the reference's (torch init order changed?)"
    sds = [{k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and "running" not in k and k != "position.pe")
            for k, v in m.state_dict().items()} for m in mods]
    ag = PR.Agent("X", sds[0], sds[1], sds[2] if len(sds) > 2 else None, **kw)
    random.seed(1)
    torch.manual_seed(7)
    if kind == "ENVDROP":
        t1, l1 = PR.rollout_envdrop(ag, penv, train_ml=True, train_rl=False, feedback="teacher")
        t2, l2 = PR.rollout_envdrop(ag, penv, train_ml=False, train_rl=True, restart=True, feedback="sample")
        l1, l2 = l1["ml_loss"], l2["rl_loss"]
        loss = l1 + l2
    else:
        fn = PR.rollout_follower if kind == "FOLLOWER" else PR.rollout_monitor
        r1 = fn(ag, penv, feedback="teacher")
        r2 = fn(ag, penv, feedback="sample", train_cl=True)
        t1, l1, t2, l2 = r1[0], r1[1], r2[0], r2[1]
        loss = l1 + l2.sum()
    loss.backward()
    assert t1 == g["traj1"] and t2 == g["traj2"]
    close(l1.detach(), g["l1"], 1e-5), close(l2.detach(), g["l2"], 1e-5)
    gn = [float(v.grad.norm()) if v.grad is not None else 0.0 for sd in sds for v in sd.values() if v.requires_grad]
    assert len(gn) == len(g["grad_norms"])
    for a, b in zip(gn, g["grad_norms"]):
        assert abs(a - b) <= 1e-4 * max(1e-3, abs(b)), (a, b)


def _eval_fixture():
    import json
    import random
    import clvln_b200  # noqa: F401
    from clvln_b200.environ import R2RBatch, make_items, make_world
    d = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eval.json")))
    w = make_world(**d["world"])
    items = make_items(w, d["items"]["n"], seed=d["items"]["seed"], instr_per_path=d["items"]["instr_per_path"])
    random.seed(0)
    return d, w, items


def test_host_evaluation_matches_reference_golden():
    """engine.Evaluation.score (host) == the reference's Evaluation.score outputs committed in tests/golden/eval.json
    (generated by oracle/make_golden.py from the unmodified src/engine/evaluator.py on a seeded synthetic world)."""
    import numpy as np
    from clvln_b200.engine import Evaluation
    from clvln_b200.environ import R2RBatch
    d, w, items = _eval_fixture()
    env = R2RBatch(w, items, batch_size=8, name="val")
    summary, per = Evaluation(env).score(d["results"])
    for k, v in d["summary"].items():
        assert np.isclose(summary[k], v, rtol=2e-6, atol=1e-6), (k, summary[k], v)
    for k in ("ndtws", "sdtws", "clss", "nav_errors", "success_path_length"):
        assert np.allclose(per[k], d["scores"][k], rtol=2e-6, atol=1e-6), k


@pytest.mark.gpu
def test_device_evaluation_matches_reference_golden():
    """The batched GPU scorer (csrc/eval.cu, Evaluation.score_device) against the same reference outputs; the tolerance
    is the fp32 storage of the distance table (the reference keeps float64 distances)."""
    import numpy as np
    import torch
    from clvln_b200 import ops
    from clvln_b200.engine import Evaluation
    from clvln_b200.environ import R2RBatch
    d, w, items = _eval_fixture()
    dev = torch.device("cuda:0")
    env = R2RBatch(w, items, batch_size=8, name="val", device=dev)
    store = ops.FeatureStore.from_world(w, dev)
    summary, m = Evaluation(env).score_device(d["results"], store)
    for k, v in d["summary"].items():
        assert np.isclose(summary[k], v, rtol=2e-6, atol=1e-6), (k, summary[k], v)
    for col, k in ((0, "nav_errors"), (1, "oracle_errors"), (3, "trajectory_lengths"), (4, "success_path_length"),
                   (5, "ndtws"), (6, "sdtws"), (7, "clss")):
        assert np.allclose(m[:, col], d["scores"][k], rtol=2e-6, atol=1e-6), k


# ---- speaker --------------------------------------------------------------------------------------------------------
def _speaker_golden():
    return torch.load(os.path.join(G, "speaker.pt"), weights_only=False)


def _speaker_world(g):
    import clvln_b200  # noqa: F401
    from clvln_b200.environ import make_world, make_items
    w = g["world"]
    world = make_world(n_scans=w["n_scans"], seed=w["seed"])
    return world, make_items(world, w["n_items"], seed=w["seed"])


def test_port_speaker_matches_reference_golden():
    """oracle/port_speaker.py on the regenerated world + weights reproduces what the REAL Speaker class produced
    (tests/golden/speaker.pt, oracle/make_golden.py): path lengths, feature checksums, eval loss / accuracies,
    beam-search scores, greedy words, gradient norms of the eval-mode loss."""
    import numpy as np
    from clvln_b200.model import SpeakerEncoder, SpeakerDecoder
    from oracle import port_env as PE, port_speaker as PS
    g = _speaker_golden()
    world, items = _speaker_world(g)
    random.seed(2020)
    torch.manual_seed(2020)
    penv = PE.R2RBatchPort(PE.WorldView(world), items, batch_size=g["world"]["B"])
    # the reference builds encoder then decoder under this seed; the product modules have the same parameter containers
    # in the same order, so the regenerated weights are the reference's
    mods = [SpeakerEncoder(2176, 512, 0.6, True, 128, 0.3), SpeakerDecoder(992, 256, 0, 512, 0.6)]
    chk = [float(p.detach().double().sum()) for m in mods for p in m.parameters()]
    assert len(chk) == len(g["w_checksum"]) and max(abs(a - b) for a, b in zip(chk, g["w_checksum"])) < 1e-6
    sds = [{k: v.detach().clone().requires_grad_(v.dtype.is_floating_point) for k, v in m.state_dict().items()} for m in mods]
    port = PS.SpeakerPort(sds[0], sds[1], max_decode=g["max_decode"])
    random.seed(1)
    for ref in g["batches"]:
        obs = penv.reset()
        assert [ob["instr_id"] for ob in obs] == ref["instr_ids"]
        (img, can), lens, _ = PS.from_shortest_path(penv, obs)
        assert lens.tolist() == ref["lengths"]
        assert abs(float(can.double().sum()) - ref["can_sum"]) < 1e-6 * max(1.0, abs(ref["can_sum"]))
        assert abs(float(img.double().sum()) - ref["img_sum"]) < 1e-6 * max(1.0, abs(ref["img_sum"]))
        insts = torch.from_numpy(np.array([ob["instr_encoding"] for ob in obs]))
        feats = ((img, can), lens)
        loss, wa, sa, _ = port.teacher_forcing(feats, insts, train=False)
        assert abs(loss - ref["loss"]) < 1e-5 * max(1.0, abs(ref["loss"])) and abs(wa - ref["word_accu"]) < 1e-9 and sa == ref["sent_accu"]
        close(port.teacher_forcing(feats, insts, train=False, for_listener=True).detach(), ref["scores"], 1e-5)
        for sd in sds:
            for v in sd.values():
                v.grad = None
        l2 = port.teacher_forcing(feats, insts, train=True, drop=None)
        l2.backward()
        assert abs(float(l2) - ref["loss_grad"]) < 1e-5 * max(1.0, abs(ref["loss_grad"]))
        names = [(0, n) for n, _ in mods[0].named_parameters()] + [(1, n) for n, _ in mods[1].named_parameters()]
        gn = [float(sds[k][n].grad.norm()) if sds[k][n].grad is not None else 0.0 for k, n in names]
        for a, b in zip(gn, ref["grad_norms"]):
            assert abs(a - b) <= 1e-4 * max(1e-3, abs(b)), (a, b)
        words, _ = port.infer_batch(feats)
        assert torch.equal(torch.from_numpy(words), ref["words"])


# ---- beam search ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["ENVDROP", "FOLLOWER", "MONITOR"])
def test_port_beam_search_matches_reference_golden(kind):
    """oracle/port_beam.py on the regenerated world + weights finds the K best listener paths the REAL agents' _dijkstra
    found (tests/golden/beam.pt, oracle/make_golden.py): same poses, actions, navigation path; scores to 1e-5."""
    import numpy as np
    from clvln_b200.model import EncoderLSTM, EnvDropDecoder, AttnDecoderLSTM, MonitorDecoder, Critic
    from oracle import port_beam as PB, port_env as PE, port_rollout as PR
    g = torch.load(os.path.join(G, "beam.pt"), weights_only=False)
    world, items = _speaker_world(g)
    random.seed(2020)
    torch.manual_seed(2020)
    penv = PE.R2RBatchPort(PE.WorldView(world), items, batch_size=g["world"]["B"])
    if kind == "ENVDROP":
        mods = [EncoderLSTM(992, 256, 512, 0, 0.5, True, 1), EnvDropDecoder(512, 0.5, 0.3, 64, 128, 2176), Critic(512, 0.5)]
        kw = dict(hidden=512, bidirectional=True, enc_layers=1, episode_len=12)
    elif kind == "FOLLOWER":
        mods = [EncoderLSTM(992, 300, 256, 0, 0.5, True, 2), AttnDecoderLSTM(256, 0.5, 2176, 2176)]
        kw = dict(hidden=256, bidirectional=True, enc_layers=2, episode_len=10)
    else:
        mods = [EncoderLSTM(992, 256, 512, 0, 0.5, False, 1), MonitorDecoder(512, 0.5, 80, [1024], 2176, 2176)]
        kw = dict(hidden=512, bidirectional=False, enc_layers=1, episode_len=10)
    ref = g[kind]
    chk = [float(p.detach().double().sum()) for m in mods for p in m.parameters()]
    assert len(chk) == len(ref["w_checksum"]) and max(abs(a - b) for a, b in zip(chk, ref["w_checksum"])) < 1e-6
    sds = [{k: v.detach().clone() for k, v in m.state_dict().items()} for m in mods]
    ag = PR.Agent(kind, sds[0], sds[1], sds[2] if len(sds) > 2 else None, **kw)
    random.seed(1)                                         # (the reference agent's constructor re-seeds `random`, base.py:28)
    with torch.no_grad():
        got = PB.dijkstra(ag, penv, g["K"], full_length=(kind == "MONITOR"))
    assert [r["instr_id"] for r in got] == [r["instr_id"] for r in ref["results"]]
    key = lambda p: (tuple(p["action"]), tuple(x[0] for x in p["trajectory"]))        # noqa: E731
    for a, b in zip(got, ref["results"]):
        assert a["dijk_path"] == b["dijk_path"]
        pa, pb = sorted(a["paths"], key=key), sorted(b["paths"], key=key)
        assert len(pa) == len(pb)
        for x, y in zip(pa, pb):
            assert [tuple(t) for t in x["trajectory"]] == y["trajectory"]
            assert x["action"] == y["action"] and x["listener_actions"] == y["listener_actions"]
            assert np.allclose(x["listener_scores"], y["listener_scores"], rtol=1e-5, atol=1e-6)
