"""CPU tests of the host logic: the C-ABI library loads and exports every symbol the header
declares; the index-table environment agrees with the oracle's dict-and-loop environment;
curriculum bookkeeping (CLR2R index map, self-paced weights) follows the reference; the N>1 data
path (minibatch sharding, SELF-PACE all-gather) works under world_size-2 gloo."""
import os
import random
import re
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_library_exports_every_declared_symbol():
    import __graft_entry__ as G
    from clvln_b200 import _lib
    G.build()
    hdr = open(os.path.join(ROOT, "include", "vln_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(vln_[a-z0-9_]+)\s*\(", hdr))
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if l.strip()}
    assert declared and declared <= exported, sorted(declared - exported)
    assert declared == set(_lib.exported_symbols()), sorted(declared ^ set(_lib.exported_symbols()))
    L = _lib.lib()                                   # loads without a GPU; no compute calls here
    assert L.vln_version() >= 100
    # the ctypes signatures follow the header: same number of parameters for every entry point
    for name, params in re.findall(r"\b(vln_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(_lib._SIGS[name][0]), (name, n, len(_lib._SIGS[name][0]))


def _world(n_items=60, seed=3):
    import clvln_b200  # noqa: F401
    from clvln_b200.environ import make_world, make_items
    w = make_world(n_scans=3, seed=seed)
    return w, make_items(w, n_items, seed=seed)


def test_obs_dict_face_matches_oracle_env():
    """reset / teacher-following steps: features, candidates, teacher, distance bit-exact vs the oracle env."""
    from clvln_b200.environ import R2RBatch
    from oracle import port_env as PE
    world, items = _world()
    random.seed(5)
    env = R2RBatch(world, items, batch_size=6)
    random.seed(5)
    penv = PE.R2RBatchPort(PE.WorldView(world), items, batch_size=6)
    obs, pobs = env.reset(), penv.reset()
    for step in range(6):
        assert len(obs) == len(pobs)
        acts = []
        for a, b in zip(obs, pobs):
            for k in ("instr_id", "scan", "viewpointId", "viewIndex", "heading", "elevation", "teacher", "instr_length"):
                assert a[k] == b[k], k
            assert np.float32(a["distance"]) == np.float32(b["distance"])
            assert np.array_equal(a["feature"], b["feature"])
            assert len(a["candidates"]) == len(b["candidates"])
            for ca, cb in zip(a["candidates"], b["candidates"]):
                assert ca["nextViewpointId"] == cb["nextViewpointId"] and ca["absViewIndex"] == cb["absViewIndex"]
                assert np.array_equal(ca["feature"], cb["feature"])
            act = -1
            for k, c in enumerate(a["candidates"]):
                if c["nextViewpointId"] == a["teacher"]:
                    act = k
            acts.append(act)
        obs, pobs = env.step(np.array(acts), obs), penv.step(np.array(acts), pobs)
    assert all(o["teacher"] == o["viewpointId"] for o in obs)          # everyone reached the goal


def test_index_face_matches_obs_face():
    from clvln_b200.environ import R2RBatch
    from clvln_b200.environ.world import heading_to_view
    world, items = _world()
    random.seed(9)
    env = R2RBatch(world, items, batch_size=8)
    ib = env.reset_index()
    obs = env.reset(restart=True)
    assert [int(x) for x in ib.vp] == [o["g"] for o in obs]
    assert [int(x) for x in ib.view] == [o["viewIndex"] for o in obs]
    assert [int(x) for x in ib.lengths] == [o["instr_length"] for o in obs]
    assert ib.tokens.shape[1] == int(ib.lengths[0]) and sorted(ib.lengths.tolist(), reverse=True) == ib.lengths.tolist()
    for i, o in enumerate(obs):
        assert np.array_equal(ib.tokens[i].numpy(), np.asarray(o["instr_encoding"])[:ib.tokens.shape[1]])
    hops = max(len(it["path"]) - 1 for it in env.batch)
    assert ib.teacher_steps == hops + 1
    ib2 = env.reset_index(restart=True)
    assert ib2 is ib


def test_sharding_reassembles_global_batch():
    from clvln_b200.environ import R2RBatch
    world, items = _world(100)
    random.seed(11)
    g = R2RBatch(world, items, batch_size=16)
    shards = []
    for r in range(4):
        random.seed(11)
        shards.append(R2RBatch(world, items, batch_size=4, rank=r, world_size=4))
    for _ in range(9):                                             # crosses a wrap-around
        st = random.getstate()
        g._next_minibatch()
        after = random.getstate()
        ids = [it["instr_id"] for it in g.batch]
        for r, e in enumerate(shards):
            random.setstate(st)
            e._next_minibatch()
            assert random.getstate() == after                      # every rank consumes the same stream
            assert [it["instr_id"] for it in e.batch] == ids[r::4]
            assert [it["instr_id"] for it in e.global_batch] == ids


@pytest.mark.parametrize("world_size", [2, 4, 8])
def test_teacher_rollout_length_is_global_under_data_parallelism(world_size):
    """Every rank must report the SAME teacher_steps for a minibatch (the maximum over the global batch): the trainers key
    their captured CUDA graphs on it, and ranks that disagreed would run warm-up / capture iterations — whose gradient
    all-reduces the replaying peers do not issue — at different steps and hang in NCCL (seen at 4 GPUs, where one rank
    had issued three iterations' worth of collectives more than the others).  Also through the staged (prefetch) path,
    and back to the local maximum when a batch is handed in explicitly."""
    from clvln_b200.environ import R2RBatch
    world, items = _world(160)
    ranks = []
    for r in range(world_size):
        random.seed(21)
        ranks.append(R2RBatch(world, items, batch_size=4, rank=r, world_size=world_size))
    local_differs = False
    for step in range(12):                                         # crosses a wrap-around
        st = random.getstate()
        got, local = [], []
        for e in ranks:
            random.setstate(st)
            if step % 3 == 2:
                e.prefetch(1)
            ib = e.reset_index()
            got.append(ib.teacher_steps)
            local.append(1 + max(e._hops_of(it) for it in e.batch))
        assert len(set(got)) == 1, got
        assert got[0] == max(local) == 1 + max(ranks[0]._hops_of(it) for it in ranks[0].global_batch)
        local_differs |= len(set(local)) > 1
    assert local_differs                                           # (the shards themselves do disagree: the case that hung)
    e = ranks[0]
    mine = list(e.batch)
    assert e.reset_index(batch=mine).teacher_steps == 1 + max(e._hops_of(it) for it in mine)


def test_curriculum_env_matches_oracle():
    from clvln_b200.environ import CLR2RBatch, split_rounds
    from oracle import port_env as PE
    world, items = _world(120)
    rounds = split_rounds(items)
    assert sum(len(v) for v in rounds.values()) == 120 and all(len(rounds[k]) for k in range(1, 6))
    random.seed(4)
    env = CLR2RBatch(world, rounds, batch_size=10, c_rate=0.8)
    random.seed(4)
    penv = PE.CLR2RBatchPort(PE.WorldView(world), rounds, batch_size=10, c_rate=0.8)
    assert np.array_equal(env.a, penv.a) and env.c == penv.c and len(env) == len(penv) == 120
    for _ in range(15):                                            # crosses the wrap-around reshuffle
        st = random.getstate()
        env._next_minibatch()
        random.setstate(st)
        penv._next_minibatch()
        assert env.cur_batch_index == penv.cur_batch_index
    ib = env.reset_index()
    assert ib.index.tolist() == env.cur_batch_index


class _FakeCLEnv:
    def __init__(self, n, seed=0):
        rs = np.random.RandomState(seed)
        self.a = rs.randint(1, 6, n).astype(np.float32)
        self.c = self.a.sum() * 0.8
        self.data = [{"instr_id": str(i)} for i in range(n)]

    def __len__(self):
        return len(self.a)

    def index(self, item):
        return int(item["instr_id"])


@pytest.mark.parametrize("func", ["linear", "log", "binary"])
def test_self_paced_weights_follow_reference_formulas(func):
    """curriculum.py:214-220, 428-448 restated independently in numpy float32."""
    from clvln_b200.engine import SelfPacedCurriculum
    env = _FakeCLEnv(500)
    sp = SelfPacedCurriculum(env, "cpu", pace_func=func, init_lamb=0.7, init_weight_ctrl=0.5, miu=0.2, interval=1,
                             strategy="epoch", burn_in=0)
    w0 = np.where(env.a <= 2, 1.0, 0.5).astype(np.float32)
    assert np.array_equal(sp.weight.numpy(), w0)
    rs = np.random.RandomState(1)
    loss = torch.from_numpy((rs.rand(500) * 1.5).astype(np.float32))
    sp.loss_for_item = loss
    sp.after_epoch(1, None)
    lamb = np.float32(0.7) + np.float32(0.2)
    assert abs(float(sp.lamb) - float(lamb)) < 1e-7
    l = loss.numpy()
    mask = l >= lamb
    w = w0.copy()
    w[mask] = 0.01
    if func == "linear":
        w[~mask] = 1 - l[~mask] / lamb
    elif func == "log":
        w[~mask] = np.log(l[~mask] + (1 - lamb)) / np.log(1 - lamb)
    else:
        w[~mask] = 1.0
    w[w < 0.01] = 0.01
    if np.dot(env.a, w) > env.c:
        w = w + env.a * (env.c - np.dot(env.a, w)) / (np.linalg.norm(env.a) ** 2)
        w[w <= 0] = 0.001
    assert np.allclose(sp.weight.numpy(), w, rtol=2e-5, atol=1e-6)


@pytest.mark.ref
def test_self_paced_update_matches_real_reference():
    from oracle import ref_loader
    if not ref_loader.reference_available():
        pytest.skip("no /root/reference")
    src = ref_loader.load_ref_agents()
    from src.engine.curriculum import SelfPacedCurriculum as RefSP
    from clvln_b200.engine import SelfPacedCurriculum
    env = _FakeCLEnv(400, 3)
    for func in ("linear", "log", "binary"):
        kw = dict(pace_func=func, init_lamb=0.3, init_weight_ctrl=0.4, miu=0.3, interval=1, strategy="epoch", burn_in=0)
        ref, mine = RefSP(env, "cpu", **kw), SelfPacedCurriculum(env, "cpu", **kw)
        assert torch.equal(ref.weight, mine.weight)
        loss = torch.rand(400, generator=torch.Generator().manual_seed(2))       # log pacing needs lambda < 1
        ref.lamb = ref.lamb + ref.stepsize
        ref.update_weight(loss.clone())
        mine.loss_for_item = loss.clone()
        mine.after_epoch(1, None)
        assert not torch.isnan(ref.weight).any()
        assert torch.equal(ref.weight, mine.weight), func          # bit-reproducible given identical losses


def _gloo_worker(rank, world_size, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    import sys
    sys.path.insert(0, ROOT)
    import clvln_b200  # noqa: F401
    from clvln_b200.environ import make_world, make_items, CLR2RBatch, split_rounds
    from clvln_b200.engine import SelfPacedCurriculum
    world = make_world(n_scans=3, seed=3)
    items = make_items(world, 120, seed=3)
    random.seed(2020)
    env = CLR2RBatch(world, split_rounds(items), batch_size=8, rank=rank, world_size=world_size)
    sp = SelfPacedCurriculum(env, "cpu", pace_func="linear", init_lamb=2.0, init_weight_ctrl=0.5, miu=2.0, interval=1,
                             burn_in=0)
    seen = []
    for it in range(3):
        ib = env.reset_index()
        seen.append(ib.index.tolist())
        sp.record(ib.index, ib.index.float() * 0.01 + it)              # a per-item "loss" every rank can verify
    g = torch.zeros(4) + rank + 1                                     # gradient averaging recipe of FlatOptimizer
    dist.all_reduce(g)
    q.put((rank, seen, sp.loss_for_item.clone(), (g / world_size).tolist(), [it["instr_id"] for it in env.global_batch]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_curriculum_allgather():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + random.randint(0, 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, seen0, l0, g0, gb0), (r1, seen1, l1, g1, gb1) = res
    assert gb0 == gb1                                                  # same global minibatch on both ranks
    assert torch.equal(l0, l1) and float(l0.abs().sum()) > 0           # replicated per-item losses after all-gather
    assert g0 == g1 == [1.5] * 4                                       # mean of per-rank gradients
    for a, b in zip(seen0, seen1):
        assert not set(a) & set(b) and len(a) == len(b) == 8           # disjoint shards of the global batch
    idx = torch.tensor(seen0[-1] + seen1[-1])
    assert torch.allclose(l0[idx], idx.float() * 0.01 + 2)


def test_dtw_cls_known_answers():
    """The reference's only known-answer vectors (doctests of src/utils/dtw.py:26-34 and src/utils/cls.py:31-39,
    3x4 grid graph with unit edges) reproduced by engine/evaluator.py's dtw_scores / cls_score."""
    import networkx as nx
    from clvln_b200.engine.evaluator import cls_score, dtw_scores
    d = dict(nx.all_pairs_dijkstra_path_length(nx.grid_graph([3, 4])))
    dist = lambda a, b: d[a][b]
    prediction = [(0, 0), (1, 0), (2, 0), (3, 0)]
    reference = [(0, 0), (1, 0), (2, 1), (3, 2)]
    dtw, ndtw, sdtw = dtw_scores(dist, prediction, reference)
    assert np.isclose(dtw, 3.0) and np.isclose(ndtw, 0.77880078307140488) and np.isclose(sdtw, 0.77880078307140488)
    assert np.isclose(dtw_scores(dist, prediction[:2], reference)[2], 0.0)
    reference = [(0, 0), (1, 0), (1, 1), (2, 1), (2, 2), (3, 2)]
    assert np.isclose(cls_score(dist, reference, reference), 1.0)
    prediction = [(0, 0), (0, 1), (1, 1), (2, 1), (3, 1), (3, 2)]
    assert np.isclose(cls_score(dist, reference, prediction), 0.81994915125863865)      # doctest call order cls(reference, prediction)
    prediction = [(0, 1), (1, 1), (2, 1), (3, 1)]
    assert np.isclose(cls_score(dist, reference, prediction), 0.44197196102702557)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the CUDA arm): one JSON line with the
    contract's keys, the reference marker, a cpu_baseline describing the run and a zero-copy e2e block."""
    import json
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--batch", "8", "--small",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1 and lines[0].startswith("{"), out.stdout[-500:]    # stdout carries the ONE json line only
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "episodes/s" and d["higher_is_better"] is True
    # kind: "reference" = the unmodified reference from oracle/_ref (oracle/build_ref.py), "port" = the oracle restatement
    from oracle import ref_loader
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_loader.reference_available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cpu"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0 and "workload" in d["config"]


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the product package (or tools/, except nothing) may import it;
    only tests/, __graft_entry__.smoke() and bench.py's CPU legs do."""
    offenders = []
    for top in ("curriculum-learning-for-vln_b200", "tools"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for fn in files:
                if fn.endswith(".py"):
                    src = open(os.path.join(dirpath, fn)).read()
                    if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M):
                        offenders.append(os.path.join(dirpath, fn))
    assert offenders == [], offenders
    entry = open(os.path.join(ROOT, "__graft_entry__.py")).read()
    build_src = entry[entry.index("def build()"):entry.index("def smoke()")]
    assert not re.search(r"(from|import)\s+oracle\b", build_src)          # build() compiles and imports the product only


def test_seen_graph_routes_match_the_reference_floyd_graph():
    """agent/beam.py's SeenGraph (dense tables over node numbers) gives the routes of the reference's FloydGraph
    (misc.py:493-541, restated in oracle/port_beam.py and pinned there through dijk_path) for the way beam search uses it:
    edges of an expanded viewpoint, one relaxation, then routes between arbitrary seen viewpoints; ties included."""
    import clvln_b200  # noqa: F401
    from clvln_b200.agent.beam import SeenGraph
    from oracle.port_beam import FloydGraph
    rng = random.Random(7)
    for trial in range(30):
        n = rng.randint(4, 14)
        names = ["v%d" % i for i in range(n)]
        nbrs = {a: rng.sample([b for b in names if b != a], rng.randint(1, min(4, n - 1))) for a in names}
        length = {}
        for a in names:
            for b in nbrs[a]:
                length.setdefault(frozenset((a, b)), rng.choice([1.0, 1.0, 2.0, 1.5, 3.25]))     # equal lengths: ties
        mine, ref = SeenGraph(), FloydGraph()
        order = names[:]
        rng.shuffle(order)
        seen = []
        for a in order + order[:3]:                                # re-expanding a viewpoint is a no-op in the search
            assert mine.seen(a) == ref.visited(a)
            if not ref.visited(a):
                for b in nbrs[a]:
                    d = length[frozenset((a, b))]
                    mine.link(a, b, d)
                    ref.add_edge(a, b, d)
                mine.relax(a)
                ref.update(a)
                seen.append(a)
            for _ in range(6):
                x, y = rng.choice(names), rng.choice(names)
                assert mine.route(x, y) == ref.path(x, y), (trial, x, y)


@pytest.mark.parametrize("kind", ["ENVDROP", "FOLLOWER", "MONITOR"])
def test_product_beam_search_host_logic_matches_oracle_on_cpu(kind):
    """agent/beam.py's search (state dictionary, tie rule, poses, SeenGraph routes, index-form visual features, the
    Self-Monitor's full-width instruction) against oracle/port_beam.py with the SAME decoder arithmetic: a stand-in agent
    whose encoder / decoder step are the oracle's CPU modules, so only the product's host-side search logic is under test
    here (the CUDA decoder step is compared on the GPU in tests/test_beam_gpu.py)."""
    import clvln_b200  # noqa: F401
    from clvln_b200.agent import beam
    from clvln_b200.environ import R2RBatch
    from clvln_b200.model import AttnDecoderLSTM, Critic, EncoderLSTM, EnvDropDecoder, MonitorDecoder
    from oracle import port_beam as PB, port_env as PE, port_modules as P, port_rollout as PR
    world, items = _world(40, seed=1)
    torch.manual_seed(5)
    if kind == "ENVDROP":
        mods = [EncoderLSTM(992, 256, 512, 0, 0.5, True, 1), EnvDropDecoder(512, 0.5, 0.3, 64, 128, 2176), Critic(512, 0.5)]
        kw = dict(hidden=512, bidirectional=True, enc_layers=1, episode_len=12)
    elif kind == "FOLLOWER":
        mods = [EncoderLSTM(992, 300, 256, 0, 0.5, True, 2), AttnDecoderLSTM(256, 0.5, 2176, 2176)]
        kw = dict(hidden=256, bidirectional=True, enc_layers=2, episode_len=10)
    else:
        mods = [EncoderLSTM(992, 256, 512, 0, 0.5, False, 1), MonitorDecoder(512, 0.5, 80, [1024], 2176, 2176)]
        kw = dict(hidden=512, bidirectional=False, enc_layers=1, episode_len=10)
    for m in mods:
        m.eval()
    sds = [{k: v.detach().clone() for k, v in m.state_dict().items()} for m in mods]
    pag = PR.Agent(kind, sds[0], sds[1], sds[2] if len(sds) > 2 else None, **kw)
    view = PE.WorldView(world)
    full = kind == "MONITOR"

    class StandIn:
        """The agent surface beam.dijkstra uses, with the oracle's arithmetic behind it."""
        device = torch.device("cpu")
        beam_full_length = full

        def __init__(self, env):
            self.env = env
            self.scratch = PE.R2RBatchPort(view, items, batch_size=env.batch_size)      # only its observe() is used

        def store_of(self, env):
            return None

        def encoder(self, tokens, lengths):
            return P.encoder_lstm(pag.enc, tokens, lengths.long(), bidirectional=pag.bi, num_layers=pag.layers, drop_ratio=0.5,
                                  drop=None)

        def beam_start_state(self, h_t):
            return h_t if kind == "ENVDROP" else torch.zeros(h_t.shape[0], 2176)

        def decode_observation(self, store, vp, view_idx, h_t, c_t, extra, ctx, ctx_mask, ended):
            self.scratch.batch = self.env.batch
            self.scratch.state = [[world.scans[int(world.vp_scan[g])], self.env._vp_name(g), v]
                                  for g, v in zip(vp.tolist(), view_idx.tolist())]
            obs = self.scratch.observe()
            logit, h_t, c_t, extra, _, _ = PB.decode_observation(pag, obs, h_t, c_t, extra, ctx, ctx_mask.dense(), ended)
            pad = torch.full((logit.shape[0], 16 - logit.shape[1]), -float("inf"))
            return torch.cat((logit, pad), 1), h_t, c_t, extra

    random.seed(2020)
    env = R2RBatch(world, items, batch_size=5)
    random.seed(2020)
    penv = PE.R2RBatchPort(view, items, batch_size=5)
    random.seed(1)
    agent = StandIn(env)
    with torch.no_grad():
        for K in (2, 4):
            got = beam.dijkstra(agent, K)
            ref = PB.dijkstra(pag, penv, K, full_length=full)
            assert [r["instr_id"] for r in got] == [r["instr_id"] for r in ref]
            key = lambda p: (tuple(p["action"]), tuple(x[0] for x in p["trajectory"]))      # noqa: E731
            for g, r in zip(got, ref):
                assert g["dijk_path"] == r["dijk_path"]
                gp, rp = sorted(g["paths"], key=key), sorted(r["paths"], key=key)
                assert len(gp) == len(rp)
                for a, b in zip(gp, rp):
                    assert [tuple(t) for t in a["trajectory"]] == [tuple(t) for t in b["trajectory"]]
                    assert a["action"] == b["action"] and a["listener_actions"] == b["listener_actions"]
                    assert np.allclose(a["listener_scores"], b["listener_scores"], rtol=1e-6, atol=1e-7)
                    # index-form features name the panorama / candidate the oracle kept as tensors
                    for (gv, vw, slot), (f, c) in zip(a["visual_feature"], b["visual_feature"]):
                        assert torch.equal(torch.from_numpy(world.table[gv].float().numpy())[:, :8], f[:, :8])
                        assert (slot == int(world.n_cand[gv])) == bool((c == 0).all())
