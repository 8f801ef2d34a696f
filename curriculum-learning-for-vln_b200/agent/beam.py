"""Beam search ("exact K best paths under the listener", src/agent/base.py:183-464) on the device-resident environment.

The search itself is the reference's: per episode a dictionary of states keyed (viewpoint, action taken there), the
unvisited state with the best accumulated listener log-probability is expanded next, an episode stops after
``max_candidates`` finished (STOP) states or when nothing is left to expand; insertion order and the first-maximum rule
of ``max`` decide ties exactly as the reference's dict iteration does.  What changes is where a step runs: the
reference teleports one simulator per episode (``newEpisode``), rebuilds the observation dicts and copies a
[B, 36, 2176] panorama to the device for every expansion (base.py:271-290); here the batch of expanded states is two
index tensors (viewpoint, view) into the HBM feature table and one call of the agent's decoder step, and the
log-probabilities come back in a single [B, 16] read per expansion instead of one ``.item()`` per candidate.
Visual features of a path are kept as table indices and only materialised when a speaker rescores the paths
(``beam_rollout``), where padded steps must be all-zero panoramas as in base.py:431-438.
"""
import math
from collections import defaultdict

import numpy as np
import torch
import torch.nn.functional as F

from .. import ops
from ..model.units import LengthMask

START = -95                                                  # action id of the start state (base.py:241)


class SeenGraph:
    """All-pairs shortest routes over the viewpoints the search has seen so far, grown one expanded viewpoint at a time
    (the role of ``utils.FloydGraph``, misc.py:493-541; it only shapes ``dijk_path``, the walk a robot would make between
    consecutive expansions).  Dense tables over node numbers in order of first appearance: ``_d[i][j]`` the best known
    length, ``_via[i][j]`` the intermediate viewpoint of that route (-1 = the direct edge).  ``relax`` is one Floyd-Warshall
    round through a newly expanded viewpoint, scanning pairs in first-appearance order with in-place symmetric updates,
    which is what decides ties between equally long routes."""
    FAR = 95959595

    def __init__(self):
        self._num, self._names = {}, []
        self._d, self._via = [], []
        self._expanded = set()

    def _node(self, name):
        i = self._num.get(name)
        if i is None:
            i = self._num[name] = len(self._names)
            self._names.append(name)
            for row in self._d:
                row.append(self.FAR)
            for row in self._via:
                row.append(-1)
            self._d.append([self.FAR] * (i + 1))
            self._via.append([-1] * (i + 1))
        return i

    def link(self, a, b, length):
        i, j = self._node(a), self._node(b)
        if length < self._d[i][j]:
            self._d[i][j] = self._d[j][i] = length
            self._via[i][j] = self._via[j][i] = -1

    def relax(self, name):
        k = self._num.get(name)
        if k is not None:
            d, via, n = self._d, self._via, len(self._names)
            for i in range(n):
                for j in range(n):
                    if i != j and d[i][k] + d[k][j] < d[i][j]:
                        d[i][j] = d[j][i] = d[i][k] + d[k][j]
                        via[i][j] = via[j][i] = k
        self._expanded.add(name)

    def seen(self, name):
        return name in self._expanded

    def route(self, a, b):
        """Viewpoints to walk through from a to b, b included (empty when a == b)."""
        if a == b:
            return []
        i, j = self._num.get(a), self._num.get(b)
        if i is None or j is None or self._via[i][j] < 0:
            return [b]
        mid = self._names[self._via[i][j]]
        return self.route(a, mid) + self.route(mid, b)


def dijkstra(agent, max_candidates, max_expansions=500):
    """base.py:183-397.  Returns the reference's result list; a path's "visual_feature" entries are
    (viewpoint index, view index, candidate slot) triples into the feature table instead of tensors."""
    env = agent.env
    ib = env.reset_index(full_length=getattr(agent, "beam_full_length", False))
    world, store = env.world, agent.store_of(env)
    B = ib.vp.shape[0]
    dev = agent.device
    vp0, view0 = ib.vp.cpu().tolist(), ib.view.cpu().tolist()
    name = env._vp_name
    n_cand, cand_vp, cand_view = world.n_cand, world.cand_vp, world.cand_view

    from ..environ.world import view_elevation, view_heading

    def here(g, view):                                       # a location reported by the environment: its own pose floats
        return (g, view, view_heading(view), view_elevation(view))

    def through(g, view):                                    # a location reached through a candidate (base.py:352-355)
        return (g, view, (view % 12) * math.pi / 6, (view // 12 - 1) * math.pi / 6)

    def pose(loc):                                           # (viewpointId, heading, elevation) as the reference records it
        return (name(loc[0]), loc[2], loc[3])

    results = [{"scan": it["scan"], "instr_id": it["instr_id"], "instr_encoding": it["instr_encoding"],
                "dijk_path": [name(vp0[i])], "paths": []} for i, it in enumerate(env.batch)]
    ctx, h_t, c_t = agent.encoder(ib.tokens, ib.lengths)
    ctx_mask = LengthMask(ib.lengths, ctx.shape[1])
    extra0 = agent.beam_start_state(h_t)
    id2state = [{(vp0[i], START): {"next_viewpoint": vp0[i], "running_state": (h_t[i], c_t[i], extra0[i]),
                                   "location": here(vp0[i], view0[i]), "from_state_id": None, "feature": None, "score": 0,
                                   "scores": [], "actions": []}} for i in range(B)]
    visited = [set() for _ in range(B)]
    finished = [set() for _ in range(B)]
    graphs = [SeenGraph() for _ in range(B)]
    ended = np.array([False] * B)
    for _ in range(max_expansions):
        pick = [max(((sid, s) for sid, s in id2state[i].items() if sid not in visited[i]), key=lambda kv: kv[1]["score"])
                if not ended[i] else next(iter(id2state[i].items())) for i in range(B)]
        tmp_ended = []
        for i, (sid, _) in enumerate(pick):
            if not ended[i]:
                visited[i].add(sid)
                if sid[1] == -1:
                    tmp_ended.append(True)
                    finished[i].add(sid)
                    if len(finished[i]) >= max_candidates:
                        ended[i] = True
                else:
                    tmp_ended.append(False)
            else:
                tmp_ended.append(True)
        h_b = torch.stack([s["running_state"][0] for _, s in pick])
        c_b = torch.stack([s["running_state"][1] for _, s in pick])
        x_b = torch.stack([s["running_state"][2] for _, s in pick])
        cur = [s["next_viewpoint"] for _, s in pick]         # "teleport": the expanded states' viewpoints and views
        views = [s["location"][1] for _, s in pick]
        vp_t = torch.tensor(cur, dtype=torch.int32, device=dev)
        view_t = torch.tensor(views, dtype=torch.int32, device=dev)
        for i, g in enumerate(cur):                          # navigation graph of what has been seen (dijk_path only)
            vn = name(g)
            if not graphs[i].seen(vn):
                for j in range(int(n_cand[g])):
                    nxt = int(cand_vp[g, j])
                    graphs[i].link(vn, name(nxt), float(world.distance(g, nxt)))
                graphs[i].relax(vn)
            results[i]["dijk_path"].extend(graphs[i].route(results[i]["dijk_path"][-1], vn))
        logits, h_b, c_b, x_b = agent.decode_observation(store, vp_t, view_t, h_b, c_b, x_b, ctx, ctx_mask, tmp_ended)
        log_probs = F.log_softmax(logits, 1).detach().cpu().numpy()
        for i, g in enumerate(cur):
            sid, state = pick[i]
            if sid[1] == -1 or ended[i]:
                continue
            nc = int(n_cand[g])
            for j in range(nc + 1):
                lp = float(log_probs[i][j])
                new_score = state["score"] + lp
                if j < nc:
                    next_id, next_vp = (g, j), int(cand_vp[g, j])
                    location = through(next_vp, int(cand_view[g, j]))
                else:
                    next_id, next_vp = (g, -1), g
                    location = here(g, views[i])
                if next_id not in id2state[i] or new_score > id2state[i][next_id]["score"]:
                    id2state[i][next_id] = {"next_viewpoint": next_vp, "location": location,
                                            "running_state": (h_b[i], c_b[i], x_b[i]), "from_state_id": sid,
                                            "feature": (g, views[i], j), "score": new_score,
                                            "scores": state["scores"] + [lp], "actions": state["actions"] + [nc + 1]}
            if len(visited[i]) == len(id2state[i]):
                ended[i] = True
        if ended.all():
            break
    for i in range(B):
        results[i]["dijk_path"].extend(graphs[i].route(results[i]["dijk_path"][-1], results[i]["dijk_path"][0]))
    for i, result in enumerate(results):
        assert len(finished[i]) <= max_candidates
        for sid in finished[i]:
            info = {"trajectory": [], "action": [], "listener_scores": id2state[i][sid]["scores"],
                    "listener_actions": id2state[i][sid]["actions"], "visual_feature": []}
            while sid[1] != START:
                state = id2state[i][sid]
                info["trajectory"].append(pose(state["location"]))
                info["action"].append(sid[1])
                info["visual_feature"].append(state["feature"])
                sid = state["from_state_id"]
            info["trajectory"].append(pose(id2state[i][sid]["location"]))
            for k in ("trajectory", "action", "visual_feature"):
                info[k] = info[k][::-1]
            result["paths"].append(info)
    return results


def path_features(agent, paths):
    """base.py:424-440: ((img_feats [P, T, 36, 2176], can_feats [P, T, 2176]), lengths) of the P paths of one episode,
    gathered from the table; steps past a path's end are all-zero, as the reference's zero-initialised tensors are."""
    store, dev = agent.store_of(agent.env), agent.device
    lengths = [len(p["visual_feature"]) for p in paths]
    P, T = len(paths), max(lengths)
    flat = [(j, k, f) for j, p in enumerate(paths) for k, f in enumerate(p["visual_feature"])]
    vp = torch.tensor([f[0] for _, _, f in flat], dtype=torch.int32, device=dev)
    view = torch.tensor([f[1] for _, _, f in flat], dtype=torch.int32, device=dev)
    slot = torch.tensor([f[2] for _, _, f in flat], dtype=torch.int32, device=dev)
    row = torch.tensor([j * T + k for j, k, _ in flat], dtype=torch.int64, device=dev)
    img = torch.zeros((P * T, ops.N_VIEWS, ops.F_DIM), device=dev)
    can = torch.zeros((P * T, ops.F_DIM), device=dev)
    img[row] = ops.gather_pano(store, vp, view)
    moves = (slot < store.n_cand[vp.long()]).unsqueeze(1)                 # the END action's feature is the all-zero row
    can[row] = ops.gather_action_feat(store, vp, view, slot, None) * moves.to(can.dtype)
    return (img.view(P, T, ops.N_VIEWS, ops.F_DIM), can.view(P, T, ops.F_DIM)), lengths


def beam_rollout(agent, speaker, beam_size):
    """base.py:399-450: the K best listener paths of every episode, each rescored by the speaker (per-word log-probs)."""
    results = dijkstra(agent, beam_size)
    eos = agent.tokenizer.word_to_index["<EOS>"]
    for result in results:
        paths = result["paths"]
        if len(paths) == 0:
            continue
        for p in paths:
            assert len(p["trajectory"]) == len(p["visual_feature"]) + 1
        features = path_features(agent, paths)
        insts = np.array([result["instr_encoding"] for _ in paths])
        seq_lengths = np.argmax(insts == eos, axis=1)
        scores = speaker.teacher_forcing(train=True, features=features, insts=torch.from_numpy(insts).to(agent.device),
                                         for_listener=True)
        scores = scores.detach().cpu().numpy()
        for j, p in enumerate(paths):
            p.pop("visual_feature")
            p["speaker_scores"] = -scores[j][:seq_lengths[j]]
    return results


def beam_search(agent, speaker, beam_size=30):
    """base.py:452-464: beam_rollout over the whole env until an instruction comes round again."""
    agent.eval()
    looped = False
    agent.results = {}
    while True:
        for traj in beam_rollout(agent, speaker, beam_size):
            if traj["instr_id"] in agent.results:
                looped = True
            else:
                agent.results[traj["instr_id"]] = traj
        if looped:
            break
