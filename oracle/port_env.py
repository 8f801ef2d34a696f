"""Oracle port of the reference's batched environment — TEST INFRASTRUCTURE ONLY.

Dict-and-loop restatement of R2RBatch / CLR2RBatch / EnvBatch (common_env.py:33-365,
curriculum_env.py:26-102) over a synthetic World description: string viewpoint ids,
an in-RAM feature dict, per-viewpoint candidate records, networkx all-pairs shortest paths.
It deliberately shares no code with the product's index-table environment; tests compare
the two, and tests/golden pins this port against the real reference run through
oracle/ref_harness.py.  Minibatching draws from Python's global ``random`` stream exactly
where the reference does (ctor shuffle :148, wrap-around reshuffle :197-200, reset_epoch :212).
"""
import math
import random

import networkx as nx
import numpy as np

ANGLE_INC = math.pi / 6.0


def angle_feat(heading, elevation, size=128):
    """ImageFeatures.make_angle_feat, misc.py:285-293."""
    return np.array([math.sin(heading), math.cos(heading), math.sin(elevation), math.cos(elevation)],
                    dtype=np.float32).repeat(size // 4)


def loc_embedding(view_index):
    """ImageFeatures.build_viewpoint_loc_embedding, misc.py:295-312."""
    emb = np.zeros((36, 128), np.float32)
    for a in range(36):
        rel = (a - view_index) % 12 + (a // 12) * 12
        emb[a] = angle_feat((rel % 12) * ANGLE_INC, (rel // 12 - 1) * ANGLE_INC)
    return emb


_LOC_EMB = [loc_embedding(v) for v in range(36)]


class WorldView:
    """String-keyed view of a World (what the reference reads from disk + simulator)."""

    def __init__(self, world):
        self.features = {}
        self.cands = {}                 # long_id -> [ {nextViewpointId, absViewIndex, normalized_heading, loc_elevation} ]
        self.graphs = {}
        table = world.table.float().cpu().numpy()
        for s, scan in enumerate(world.scans):
            G = nx.Graph()
            for (u, v), w in world.edge_len[s].items():
                G.add_edge(world.vp_names[s][u], world.vp_names[s][v], weight=w)
            self.graphs[scan] = G
            o = int(world.scan_off[s])
            for j, name in enumerate(world.vp_names[s]):
                g = o + j
                lid = f"{scan}_{name}"
                self.features[lid] = table[g]
                self.cands[lid] = [dict(
                    nextViewpointId=world.vp_names[s][int(world.cand_vp[g, k]) - o],
                    absViewIndex=int(world.cand_view[g, k]),
                    normalized_heading=float(world.cand_nheading[g, k]),
                    loc_elevation=float(world.cand_elev[g, k])) for k in range(int(world.n_cand[g]))]


class R2RBatchPort:
    def __init__(self, view, items, batch_size=100, name="train"):
        self.view = view
        self.data = [dict(it) for it in items]
        self.name = name
        self.scans = set(it["scan"] for it in self.data)
        random.shuffle(self.data)                                   # common_env.py:148
        self.ix = 0
        self.batch_size = batch_size
        self.paths, self.distances = {}, {}
        for scan in self.scans:                                     # common_env.py:164-181
            G = view.graphs[scan]
            self.paths[scan] = dict(nx.all_pairs_dijkstra_path(G))
            self.distances[scan] = dict(nx.all_pairs_dijkstra_path_length(G))
        self.state = []                                             # per episode [scan, vp, viewIndex]

    def size(self):
        return len(self.data)

    def _next_minibatch(self, sort=True):                           # common_env.py:183-207
        batch = self.data[self.ix: self.ix + self.batch_size]
        if len(batch) < self.batch_size:
            random.shuffle(self.data)
            self.ix = self.batch_size - len(batch)
            batch += self.data[:self.ix]
        else:
            self.ix += self.batch_size
        if sort:
            batch = sorted(batch, key=lambda it: it["instr_length"], reverse=True)
        self.batch = batch

    def reset_epoch(self, shuffle=False):                           # common_env.py:209-214
        if shuffle:
            random.shuffle(self.data)
        self.ix = 0

    def _candidates(self, scan, vp, view_index, img):               # common_env.py:275-295 (buffered path)
        base = (view_index % 12) * ANGLE_INC
        out = []
        for rec in self.view.cands[f"{scan}_{vp}"]:
            c = dict(rec)
            c["scanId"] = scan
            c["loc_heading"] = c.pop("normalized_heading") - base
            c["feature"] = np.concatenate(
                (img[c["absViewIndex"]], angle_feat(c["loc_heading"], c["loc_elevation"])), -1)
            out.append(c)
        return out

    def observe(self):                                              # common_env.py:299-330
        obs = []
        for i, (scan, vp, vi) in enumerate(self.state):
            item = self.batch[i]
            img = self.view.features[f"{scan}_{vp}"]
            goal = item["path"][-1]
            teacher = goal if vp == goal else self.paths[scan][vp][goal][1]
            obs.append({
                "instr_id": item["instr_id"], "scan": scan, "viewpointId": vp, "viewIndex": vi,
                "heading": (vi % 12) * ANGLE_INC, "elevation": (vi // 12 - 1) * ANGLE_INC,
                "feature": np.concatenate((img, _LOC_EMB[vi]), -1),
                "candidates": self._candidates(scan, vp, vi, img),
                "instructions": item["instructions"], "teacher": teacher, "path_id": item["path_id"],
                "instr_encoding": item["instr_encoding"], "instr_length": item["instr_length"],
                "distance": self.distances[scan][vp][goal],
            })
        return obs

    def reset(self, batch=None, inject=False, restart=False, **kw):  # common_env.py:332-348
        if not restart:
            if batch is None:
                self._next_minibatch(**kw)
            elif inject:
                self._next_minibatch(**kw)
                self.batch[:len(batch)] = batch
            else:
                self.batch = batch
        self.state = [[it["scan"], it["path"][0], 12 + int(round(it["heading"] / ANGLE_INC)) % 12]
                      for it in self.batch]
        return self.observe()

    def step(self, actions, obs, traj=None):                        # common_env.py:91-110, 350-353
        for i, a in enumerate(np.asarray(actions).tolist()):
            if a == -1:
                continue
            cand = obs[i]["candidates"][a]
            self.state[i][1] = cand["nextViewpointId"]
            self.state[i][2] = cand["absViewIndex"]
            if traj is not None:
                vi = self.state[i][2]
                traj[i]["path"].append((self.state[i][1], (vi % 12) * ANGLE_INC, (vi // 12 - 1) * ANGLE_INC))
        return self.observe()


class CLR2RBatchPort(R2RBatchPort):
    """curriculum_env.py:26-102: rounds 1..5 concatenated; a[i] = round of item i in
    round-major order; c = sum(a) * c_rate."""

    def __init__(self, view, rounds, batch_size=100, c_rate=0.8):
        items = [it for k in range(1, 6) for it in rounds[k]]
        super().__init__(view, items, batch_size, "train")
        self.a = np.zeros(len(self.data), np.float32)
        self.item2idx = {}
        for k in range(1, 6):
            for it in rounds[k]:
                i = len(self.item2idx)
                self.item2idx[it["instr_id"]] = i
                self.a[i] = k
        self.c = self.a.sum() * c_rate

    def __len__(self):
        return len(self.data)

    def index(self, item):
        return self.item2idx[item["instr_id"]]

    @property
    def cur_batch_index(self):
        return [self.item2idx[it["instr_id"]] for it in self.batch]
