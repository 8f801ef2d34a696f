"""Oracle port of the agents' beam search — TEST INFRASTRUCTURE ONLY.

Restates ``BasicR2RAgent._dijkstra`` (src/agent/base.py:183-397: the exact-K-best search over (viewpoint, action) states
under the listener's log-probabilities), the agents' ``running_state`` / ``decode_obervation`` hooks (envdrop.py:280-297,
follower.py:175-198, monitor.py:201-225), ``FloydGraph`` (src/utils/misc.py:493-541) and ``beam_rollout``'s speaker
rescoring (base.py:399-450), on oracle/port_modules.py + oracle/port_speaker.py and the obs-dict environment of
oracle/port_env.py ("teleporting" = writing the episode's [scan, viewpoint, viewIndex] state and observing again, which is
what ``sims[i].newEpisode`` + ``env.observe()`` do in the reference).  Pinned by tests/_ref_check_beam.py against the
unmodified reference agents' ``_dijkstra`` in this container.
"""
import math
from collections import defaultdict

import numpy as np
import torch
import torch.nn.functional as F

from . import port_modules as P
from . import port_rollout as PR


class FloydGraph:                                            # misc.py:493-541
    def __init__(self):
        self._dis = defaultdict(lambda: defaultdict(lambda: 95959595))
        self._point = defaultdict(lambda: defaultdict(lambda: ""))
        self._visited = set()

    def add_edge(self, x, y, dis):
        if dis < self._dis[x][y]:
            self._dis[x][y] = dis
            self._dis[y][x] = dis
            self._point[x][y] = ""
            self._point[y][x] = ""

    def update(self, k):
        for x in self._dis:
            for y in self._dis:
                if x != y:
                    if self._dis[x][k] + self._dis[k][y] < self._dis[x][y]:
                        self._dis[x][y] = self._dis[x][k] + self._dis[k][y]
                        self._dis[y][x] = self._dis[x][y]
                        self._point[x][y] = k
                        self._point[y][x] = k
        self._visited.add(k)

    def visited(self, k):
        return k in self._visited

    def path(self, x, y):
        if x == y:
            return []
        if self._point[x][y] == "":
            return [y]
        k = self._point[x][y]
        return self.path(x, k) + self.path(k, y)


def decode_observation(ag, obs, h_t, c_t, extra, ctx, seq_mask, ended):
    """One listener step for the expanded states: (masked logits, h_t, c_t, extra, img [B,36,F], cands [B,C,F])."""
    dev = ag.device
    img = PR.pano_tensor(obs, dev)
    cands, lens = PR.cand_tensor(obs, dev)
    cmask = PR.length_mask(lens, dev)
    if ag.kind == "ENVDROP":
        logit, (h_t, c_t), extra, _ = P.envdrop_decoder(ag.dec, PR.pose_tensor(obs, dev), img, cands, extra, c_t, ctx, seq_mask,
                                                        drop_ratio=ag.p, feat_drop_ratio=ag.pf, drop=None)
        return logit.masked_fill(cmask, -float("inf")), h_t, c_t, extra, img, cands
    if ag.kind == "FOLLOWER":
        logit, (h_t, c_t), _ = P.follower_decoder(ag.dec, img, extra, cands, h_t, c_t, ctx, seq_mask, drop_ratio=ag.p, drop=None)
    else:
        (logit, _), (h_t, c_t), _ = P.monitor_decoder(ag.dec, extra, cands, h_t, c_t, ctx, seq_mask, cmask, training=False)
    logit = logit.masked_fill(cmask, -float("inf"))
    a = logit.max(1)[1].detach().cpu().numpy().copy()
    for i, nid in enumerate(a):
        if nid == len(obs[i]["candidates"]) or nid == -1 or ended[i]:
            a[i] = -1
    extra = cands[np.arange(len(obs)), np.maximum(a, 0), :].detach()
    return logit, h_t, c_t, extra, img, cands


def dijkstra(ag, env, max_candidates, full_length=False):
    """base.py:183-397 on (port agent ``ag``, obs-dict env).  ``full_length``: the Self-Monitor's full-width instruction."""
    dev = ag.device
    obs = env.reset()
    B = len(obs)
    results = [{"scan": ob["scan"], "instr_id": ob["instr_id"], "instr_encoding": ob["instr_encoding"],
                "dijk_path": [ob["viewpointId"]], "paths": []} for ob in obs]
    seq, seq_mask, lengths = PR.instr_tensors(obs, full_length, dev)
    ctx, h_t, c_t = P.encoder_lstm(ag.enc, seq, lengths, bidirectional=ag.bi, num_layers=ag.layers, drop_ratio=ag.p, drop=None)
    extra0 = h_t if ag.kind == "ENVDROP" else torch.zeros(B, 2176, device=dev)
    sid_of = lambda vp, a: "%s_%s" % (vp, str(a))                         # noqa: E731
    id2state = [{sid_of(ob["viewpointId"], -95): {
        "next_viewpoint": ob["viewpointId"], "running_state": (h_t[i], c_t[i], extra0[i]),
        "location": (ob["viewpointId"], ob["heading"], ob["elevation"]), "from_state_id": None, "feature": None,
        "score": 0, "scores": [], "actions": []}} for i, ob in enumerate(obs)]
    visited = [set() for _ in range(B)]
    finished = [set() for _ in range(B)]
    graphs = [FloydGraph() for _ in range(B)]
    ended = np.array([False] * B)
    for _ in range(500):
        pick = [max(((sid, s) for sid, s in id2state[i].items() if sid not in visited[i]), key=lambda it: it[1]["score"])
                if not ended[i] else next(iter(id2state[i].items())) for i in range(B)]
        tmp_ended = []
        for i, (sid, _) in enumerate(pick):
            if not ended[i]:
                action = int(sid.rsplit("_", 1)[1])
                visited[i].add(sid)
                if action == -1:
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
        for i, (_, s) in enumerate(pick):                    # newEpisode(scan, next_viewpoint, heading, elevation)
            _, heading, elevation = s["location"]
            view = (1 + int(round(elevation / (math.pi / 6)))) * 12 + int(round(heading / (math.pi / 6))) % 12
            env.state[i] = [results[i]["scan"], s["next_viewpoint"], view]
        obs = env.observe()
        for i, ob in enumerate(obs):
            vp = ob["viewpointId"]
            if not graphs[i].visited(vp):
                for c in ob["candidates"]:
                    graphs[i].add_edge(vp, c["nextViewpointId"], env.distances[ob["scan"]][vp][c["nextViewpointId"]])
                graphs[i].update(vp)
            results[i]["dijk_path"].extend(graphs[i].path(results[i]["dijk_path"][-1], vp))
        logits, h_b, c_b, x_b, f_t, cand_feat = decode_observation(ag, obs, h_b, c_b, x_b, ctx, seq_mask, tmp_ended)
        log_probs = F.log_softmax(logits, 1)
        for i, ob in enumerate(obs):
            cur_vp, cand = ob["viewpointId"], ob["candidates"]
            cur_id, cur = pick[i]
            from_action = int(cur_id.rsplit("_", 1)[1])
            assert cur_vp == cur["next_viewpoint"]
            if from_action == -1 or ended[i]:
                continue
            for j in range(len(cand) + 1):
                lp = log_probs[i][j].detach().cpu().item()
                new_score = cur["score"] + lp
                if j < len(cand):
                    next_id, next_vp = sid_of(cur_vp, j), cand[j]["nextViewpointId"]
                    trg = cand[j]["absViewIndex"]
                    location = (next_vp, (trg % 12) * math.pi / 6, (trg // 12 - 1) * math.pi / 6)
                else:
                    next_id, next_vp = sid_of(cur_vp, -1), cur_vp
                    location = (cur_vp, ob["heading"], ob["elevation"])
                if next_id not in id2state[i] or new_score > id2state[i][next_id]["score"]:
                    id2state[i][next_id] = {"next_viewpoint": next_vp, "location": location,
                                            "running_state": (h_b[i], c_b[i], x_b[i]), "from_state_id": cur_id,
                                            "feature": (f_t[i].detach().cpu(), cand_feat[i][j].detach().cpu()),
                                            "score": new_score, "scores": cur["scores"] + [lp],
                                            "actions": cur["actions"] + [len(cand) + 1]}
            if len(visited[i]) == len(id2state[i]):
                ended[i] = True
        if ended.all():
            break
    for i in range(B):
        results[i]["dijk_path"].extend(graphs[i].path(results[i]["dijk_path"][-1], results[i]["dijk_path"][0]))
    for i, result in enumerate(results):
        for sid in finished[i]:
            info = {"trajectory": [], "action": [], "listener_scores": id2state[i][sid]["scores"],
                    "listener_actions": id2state[i][sid]["actions"], "visual_feature": []}
            action = int(sid.rsplit("_", 1)[1])
            while action != -95:
                st = id2state[i][sid]
                info["trajectory"].append(st["location"])
                info["action"].append(action)
                info["visual_feature"].append(st["feature"])
                sid = st["from_state_id"]
                action = int(sid.rsplit("_", 1)[1])
            info["trajectory"].append(id2state[i][sid]["location"])
            for k in ("trajectory", "action", "visual_feature"):
                info[k] = info[k][::-1]
            result["paths"].append(info)
    return results


def speaker_scores(spk, result, eos=2, drop=None):
    """base.py:419-449 for one episode's result: per-path negative per-word CE of the instruction under the speaker
    (``spk``: port_speaker.SpeakerPort), cut at the instruction's length."""
    paths = result["paths"]
    lengths = [len(p["visual_feature"]) for p in paths]
    T, n = max(lengths), len(paths)
    img = torch.zeros(n, T, 36, 2176)
    can = torch.zeros(n, T, 2176)
    for j, p in enumerate(paths):
        for k, (f, c) in enumerate(p["visual_feature"]):
            img[j][k] = f
            can[j][k] = c
    insts = np.array([result["instr_encoding"] for _ in range(n)])
    seq_lengths = np.argmax(insts == eos, axis=1)
    sc = spk.teacher_forcing(((img, can), lengths), torch.from_numpy(insts), train=True, for_listener=True, drop=drop)
    return [-sc[j].detach().numpy()[:seq_lengths[j]] for j in range(n)]
