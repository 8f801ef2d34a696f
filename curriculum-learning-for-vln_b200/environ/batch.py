"""Batched R2R environment over the World's index tables (the stub for MatterSim + dataset).

Mirrors the protocol the reference's agents and trainers use (common_env.py:117-365,
curriculum_env.py:26-102): ``reset(restart=, batch=, inject=)``, ``observe()``,
``step(actions, obs, traj)``, ``reset_epoch(shuffle)``, ``batch``, ``size()``, ``distances``,
and for curricula ``a``, ``c``, ``index(item)``, ``cur_batch_index``, ``len(env)``.

Minibatch order is the reference's, drawn from Python's GLOBAL ``random`` stream at the same
three places (ctor shuffle common_env.py:148, wrap-around reshuffle :197-200, reset_epoch :212),
followed by the same stable sort by instruction length (:204-205) — bit-exact by construction
and pinned by tests/test_ordering.py.

Two faces on one state:
  * index face (what the kernels consume): ``reset_index()`` -> IndexBatch of device tensors
    (tokens, lengths, start viewpoint / view, goal, dataset index);  the simulator transition
    itself is the ``vln_env_step`` kernel, so a rollout never returns to the host;
  * obs-dict face (drop-in / evaluation / trajectories): lists of dicts with the reference's
    keys, assembled on the host from the same tables.

Data parallelism: every rank builds the same env from the same seed and draws the
same GLOBAL minibatch of world_size*batch_size items; after the sort rank r keeps rows
r::world_size, so the per-rank batch is a reference-sized, length-balanced batch and no
communication is needed.
"""
import math
import random
from dataclasses import dataclass

import numpy as np
import torch

from .world import ANGLE_INC, N_VIEWS, heading_to_view, static_loc4, view_elevation, view_heading


@dataclass
class IndexBatch:
    tokens: torch.Tensor        # int64 [B, L']  (L' = longest instruction in the batch, or max_len)
    lengths: torch.Tensor       # int32 [B] on device
    lengths_cpu: torch.Tensor   # int64 [B] on host (EncoderLSTM.forward takes host lengths, units.py:48)
    vp: torch.Tensor            # int32 [B] start viewpoint (global index)
    view: torch.Tensor          # int32 [B] start view index
    goal: torch.Tensor          # int32 [B]
    index: torch.Tensor         # int64 [B] dataset index of each episode (curriculum weights)
    teacher_steps: int          # host-known length of the teacher-forced rollout (max hops + 1 STOP)
    h2d_bytes: int


class R2RBatch:
    def __init__(self, world, items, batch_size=100, name="train", device=None, max_len=80,
                 rank=0, world_size=1):
        self.world = world
        self.data = [dict(it) for it in items]
        self.name = name
        self.scans = set(it["scan"] for it in self.data)
        self.splits = [name]
        random.shuffle(self.data)                                    # common_env.py:148
        self.ix = 0
        self.batch_size = batch_size
        self.rank, self.world_size = rank, world_size
        self.device = torch.device(device) if device is not None else None
        self.max_len = max_len
        self.feature_size = 2048
        self.batch = None
        self.global_batch = None
        self._state = None          # host state of the obs-dict face: list of [g, view]
        self._loc4 = static_loc4()
        self._name2g = None
        self.distances = _DistanceView(world)
        self._staged = []           # IndexBatches made ahead of time by prefetch()

    # ---- dataset iteration (bit-exact with the reference) ------------------------------------
    def size(self):
        return len(self.data)

    def _next_minibatch(self, tile_one=False, sort=True, **_):
        n = self.batch_size * self.world_size
        if tile_one:                                                 # common_env.py:189-194
            batch = [self.data[self.ix]] * n
            self.ix += 1
            if self.ix >= len(self.data):
                random.shuffle(self.data)
                self.ix -= len(self.data)
        else:
            batch = self.data[self.ix: self.ix + n]
            if len(batch) < n:
                random.shuffle(self.data)
                self.ix = n - len(batch)
                batch += self.data[:self.ix]
            else:
                self.ix += n
        if sort and "instr_length" in batch[0]:
            batch = sorted(batch, key=lambda it: it["instr_length"], reverse=True)
        self.global_batch = batch
        self.batch = batch[self.rank::self.world_size] if self.world_size > 1 else batch
        self._batch_is_shard = True

    def reset_epoch(self, shuffle=False):
        if shuffle:
            random.shuffle(self.data)
        self.ix = 0

    def _select(self, batch=None, inject=False, restart=False, **kw):
        if restart:
            return
        if batch is None:
            self._next_minibatch(**kw)
        elif inject:
            self._next_minibatch(**kw)
            self.batch[:len(batch)] = batch
            self._batch_is_shard = False
        else:
            self.batch = batch
            self._batch_is_shard = False

    # ---- index face -----------------------------------------------------------------------------
    def reset_index(self, batch=None, inject=False, restart=False, full_length=False, **kw):
        """Start new episodes and return their index tensors on the device (pinned staging,
        one async H2D copy per field)."""
        if restart and getattr(self, "_last_ib", None) is not None and batch is None:
            return self._last_ib                      # same episodes again (reset(restart=True))
        if self._staged and batch is None:
            self.batch, ib = self._staged.pop(0)
            self._last_ib = ib
            return ib
        self._select(batch, inject, restart, **kw)
        ib = self._make_index_batch(full_length)
        self._last_ib = ib
        return ib

    def prefetch(self, n, full_length=False):
        """Draw the next n minibatches now and stage their index tensors in HBM (benchmarks that
        want inputs resident before the timed region; order is unchanged)."""
        for _ in range(n):
            self._next_minibatch()
            self._staged.append((self.batch, self._make_index_batch(full_length)))

    def _hops_of(self, item):
        """Hops of the item's shortest path (start -> goal), cached on the item."""
        n = item.get("_hops")
        if n is None:
            w = self.world
            cur, goal, n = int(item["path_g"][0]), int(item["path_g"][-1]), 0
            while cur != goal:
                cur = w.hop(cur, goal)
                n += 1
            item["_hops"] = n
        return n

    def _make_index_batch(self, full_length=False):
        b = self.batch
        w = self.world
        lengths = np.array([it["instr_length"] for it in b], np.int64)
        L = self.max_len if full_length else int(lengths[0] if len(lengths) else 0)
        L = max(L, int(lengths.max()))
        tokens = np.stack([np.asarray(it["instr_encoding"])[:L] for it in b]).astype(np.int64)
        vp = np.array([it["path_g"][0] for it in b], np.int32)
        goal = np.array([it["path_g"][-1] for it in b], np.int32)
        view = np.array([heading_to_view(it["heading"]) for it in b], np.int32)
        index = np.array([self.index(it) for it in b], np.int64) if hasattr(self, "item2idx") else \
            np.zeros(len(b), np.int64)
        # Length of the teacher-forced rollout.  Under data parallelism it is taken over the GLOBAL minibatch (every rank
        # draws the same one and keeps rows rank::world), not over this rank's shard: the trainers key their captured CUDA
        # graphs on it, and a rank that met a new length would run warm-up + capture iterations — with their gradient
        # all-reduces — while its peers replay, i.e. the ranks would issue different numbers of collectives and hang.
        # Rows that reach their goal earlier are masked exactly as before, so losses are unchanged.
        rows = self.global_batch if (self.world_size > 1 and getattr(self, "_batch_is_shard", False)) else b
        hops = max((self._hops_of(it) for it in rows), default=0)
        dev = self.device
        nb = 0

        def put(a, dt):
            nonlocal nb
            t = torch.from_numpy(a)
            nb += t.numel() * t.element_size()
            if dev is not None and dev.type == "cuda":
                return t.pin_memory().to(dev, dtype=dt, non_blocking=True)
            return t.to(dt)
        self._state = [[int(s), int(v)] for s, v in zip(vp, view)]
        return IndexBatch(tokens=put(tokens, torch.int64), lengths=put(lengths.astype(np.int32), torch.int32),
                          lengths_cpu=torch.from_numpy(lengths), vp=put(vp, torch.int32),
                          view=put(view, torch.int32), goal=put(goal, torch.int32),
                          index=put(index, torch.int64), teacher_steps=hops + 1, h2d_bytes=nb)

    # ---- obs-dict face --------------------------------------------------------------------------
    def _g_of(self, scan, name):
        if self._name2g is None:
            self._name2g = {}
            for s, sc in enumerate(self.world.scans):
                o = int(self.world.scan_off[s])
                for j, nm in enumerate(self.world.vp_names[s]):
                    self._name2g[(sc, nm)] = o + j
        return self._name2g[(scan, name)]

    def _vp_name(self, g):
        w = self.world
        return w.vp_names[int(w.vp_scan[g])][int(w.vp_local[g])]

    def make_candidate(self, g, view_index, img=None):
        """common_env.py:225-297 on tables: candidate dicts in sweep order."""
        w = self.world
        if img is None:
            img = w.table[g].float().numpy()
        out = []
        k = view_index % 12
        for j in range(int(w.n_cand[g])):
            av = int(w.cand_view[g, j])
            ang = np.repeat(w.cand_ang4[g, j, k], 32)
            out.append({
                "scanId": w.scans[int(w.vp_scan[g])], "nextViewpointId": self._vp_name(int(w.cand_vp[g, j])),
                "absViewIndex": av, "loc_heading": float(w.cand_nheading[g, j]) - k * ANGLE_INC,
                "loc_elevation": float(w.cand_elev[g, j]),
                "feature": np.concatenate((img[av], ang), -1), "next_g": int(w.cand_vp[g, j]),
            })
        return out

    def observe(self):
        w = self.world
        obs = []
        for i, (g, vi) in enumerate(self._state):
            item = self.batch[i]
            goal = item["path_g"][-1]
            img = w.table[g].float().numpy()
            teacher_g = goal if g == goal else w.hop(g, goal)
            ob = {
                "instr_id": item["instr_id"], "scan": item["scan"], "viewpointId": self._vp_name(g),
                "viewIndex": vi, "heading": view_heading(vi), "elevation": view_elevation(vi),
                "feature": np.concatenate((img, np.repeat(self._loc4[vi], 32, axis=1)), -1),
                "candidates": self.make_candidate(g, vi, img),
                "instructions": item["instructions"], "teacher": self._vp_name(teacher_g),
                "path_id": item["path_id"], "distance": float(w.distance(g, goal)), "g": g,
            }
            if "instr_encoding" in item:
                ob["instr_encoding"] = item["instr_encoding"]
                ob["instr_length"] = item["instr_length"]
            obs.append(ob)
        return obs

    def reset(self, batch=None, inject=False, restart=False, **kw):
        self._select(batch, inject, restart, **kw)
        self._state = [[it["path_g"][0], heading_to_view(it["heading"])] for it in self.batch]
        return self.observe()

    def step(self, actions, obs, traj=None):
        for i, a in enumerate(np.asarray(actions).tolist()):
            if a == -1:
                continue
            cand = obs[i]["candidates"][a]
            self._state[i] = [cand["next_g"], cand["absViewIndex"]]
            if traj is not None:
                vi = cand["absViewIndex"]
                traj[i]["path"].append((cand["nextViewpointId"], view_heading(vi), view_elevation(vi)))
        return self.observe()

    def traj_from_index(self, vp_hist, view_hist):
        """Trajectories in the reference's result format from a device rollout's [T+1, B] state
        history (host arrays): consecutive duplicates (= no move) are dropped."""
        out = []
        for i, item in enumerate(self.batch):
            path = []
            for t in range(vp_hist.shape[0]):
                g, vi = int(vp_hist[t, i]), int(view_hist[t, i])
                if t and g == int(vp_hist[t - 1, i]) and vi == int(view_hist[t - 1, i]):
                    continue
                path.append((self._vp_name(g), view_heading(vi), view_elevation(vi)))
            out.append({"instr_id": item["instr_id"], "path": path})
        return out

    def get_statistics(self):
        n = max(1, len(self.data))
        return {"length": sum(it["instr_length"] for it in self.data) / n,
                "path": sum(it["distance"] for it in self.data) / n}


class _DistanceView:
    """env.distances[scan][vp_a][vp_b] as the evaluator reads it (evaluator.py:56-70)."""

    def __init__(self, world):
        self.w = world

    def __getitem__(self, scan):
        w = self.w
        s = w.scans.index(scan)
        o = int(w.scan_off[s])
        names = {n: o + j for j, n in enumerate(w.vp_names[s])}

        class _Row:
            def __getitem__(_, a):
                class _Col:
                    def __getitem__(__, b):
                        return float(w.distance(names[a], names[b]))
                return _Col()
        return _Row()


class CLR2RBatch(R2RBatch):
    """Curriculum dataset (curriculum_env.py:26-102): rounds 1..5 concatenated; index order is
    round-major; a[i] = round number (difficulty); c = sum(a) * c_rate."""

    def __init__(self, world, rounds, batch_size=100, c_rate=0.8, device=None, max_len=80, rank=0,
                 world_size=1):
        items = [it for k in range(1, 6) for it in rounds[k]]
        self.c_rate = c_rate
        self.curriculum_data = {f"round_{k}": list(rounds[k]) for k in range(1, 6)}
        super().__init__(world, items, batch_size, "train", device, max_len, rank, world_size)
        self.splits = [f"train_round[{k}]_v3" for k in range(1, 6)]
        self.a = np.zeros(len(self.data), np.float32)
        self.item2idx = {}
        for key, data in self.curriculum_data.items():
            for it in data:
                i = len(self.item2idx)
                self.item2idx[it["instr_id"]] = i
                self.a[i] = int(key[-1])
        self.c = self.a.sum() * self.c_rate

    def __len__(self):
        return len(self.data)

    def index(self, item):
        return self.item2idx[item["instr_id"]]

    @property
    def cur_batch_index(self):
        return [self.item2idx[it["instr_id"]] for it in self.batch]


def split_rounds(items, sizes=(1037, 1415, 4897, 4593, 2097)):
    """Deal synthetic items into 5 difficulty rounds with the CLR2R proportions;
    round = rank of the path length (hops), ties by instr_id — a stand-in for the offline
    room-count difficulty of the real CLR2R split."""
    order = sorted(items, key=lambda it: (len(it["path"]), it["instr_id"]))
    tot = float(sum(sizes))
    rounds, a = {}, 0
    for k, s in enumerate(sizes, 1):
        b = len(order) if k == len(sizes) else a + int(round(len(order) * s / tot))
        rounds[k] = order[a:b]
        a = b
    return rounds


__all__ = ["R2RBatch", "CLR2RBatch", "IndexBatch", "split_rounds"]
