"""Synthetic Matterport-shaped world: the stand-in for the simulator + dataset.

The reference walks a C++ simulator (MatterSim) over per-scan connectivity graphs
and looks features up in an in-RAM dict (common_env.py:33-110, 225-297).  There is
no simulator, dataset or network here, so the world is generated (SURVEY.md §8d
"Synthetic inputs") and held as flat index tables that live in HBM:

  table      bf16 [n_vp, 36, 2048]    panoramic ResNet-152-shaped features
  cand_vp    i32  [n_vp, CMAX]        neighbour viewpoint (global index), -1 = none
  cand_view  i32  [n_vp, CMAX]        absViewIndex the neighbour is seen in
  cand_ang4  f32  [n_vp, CMAX, 12, 4] sin/cos of (normalized_heading - base_heading[k]), elevation
  n_cand     i32  [n_vp]
  next_hop   i32  [sum n_s^2]         per-scan all-pairs "next viewpoint on the shortest path" (global idx)
  dist       f32  [sum n_s^2]         per-scan all-pairs shortest distance (metres)
  sq_off     i64  [n_vp]              offset of row (scan(vp), local(vp), 0) in next_hop / dist
  vp_local   i32  [n_vp]

plus the description the oracle's obs-dict environment consumes (string ids,
adjacency dicts) so that both faces are built from the same world.
"""
import math
import random as _pyrandom
from dataclasses import dataclass, field

import numpy as np
import torch

ANGLE_INC = math.pi / 6.0          # misc.py `angle_inc` (30 degrees)
N_VIEWS = 36
IMG_DIM = 2048
ANGLE_DIM = 128
CMAX = 15                          # max navigable neighbours; +1 END slot => 16 logits


def angle_feat4(heading: float, elevation: float):
    """The 4 distinct values of make_angle_feat (misc.py:285-293); each is repeated 32x."""
    return (math.sin(heading), math.cos(heading), math.sin(elevation), math.cos(elevation))


def static_loc4():
    """[36 cur-view, 36 abs-view, 4]: build_viewpoint_loc_embedding (misc.py:295-312),
    de-duplicated (the 128-d embedding is these 4 values .repeat(32))."""
    out = np.zeros((N_VIEWS, N_VIEWS, 4), np.float32)
    for cur in range(N_VIEWS):
        for ab in range(N_VIEWS):
            rel = (ab - cur) % 12 + (ab // 12) * 12
            out[cur, ab] = angle_feat4((rel % 12) * ANGLE_INC, (rel // 12 - 1) * ANGLE_INC)
    return out


def pose4():
    """[36, 4]: angle feature of the agent's own pose per discretised viewIndex
    (envdrop.py:76-78 with the simulator's heading/elevation grid)."""
    out = np.zeros((N_VIEWS, 4), np.float32)
    for v in range(N_VIEWS):
        out[v] = angle_feat4(view_heading(v), view_elevation(v))
    return out


def view_heading(view_idx: int) -> float:
    return (view_idx % 12) * ANGLE_INC


def view_elevation(view_idx: int) -> float:
    return (view_idx // 12 - 1) * ANGLE_INC


def heading_to_view(heading: float) -> int:
    """newEpisode(scan, vp, heading, 0) with discretised angles: elevation row 1, heading
    snapped to the 30-degree grid."""
    return 12 + int(round(heading / ANGLE_INC)) % 12


@dataclass
class World:
    scans: list
    vp_names: list                       # per scan: list of viewpoint id strings
    scan_off: np.ndarray                 # i64 [n_scans+1]
    cand_vp: np.ndarray
    cand_view: np.ndarray
    cand_nheading: np.ndarray            # f64 [n_vp, CMAX] normalized_heading
    cand_elev: np.ndarray                # f64 [n_vp, CMAX] loc_elevation
    cand_rel: np.ndarray                 # f64 [n_vp, CMAX] rel_heading inside the view (nheading = view heading + rel)
    n_cand: np.ndarray
    edge_len: list                       # per scan: dict {(u_local, v_local): metres}, u < v
    table: torch.Tensor = None           # bf16 [n_vp, 36, 2048]
    # derived (build_routes)
    next_hop: np.ndarray = None
    dist: np.ndarray = None
    sq_off: np.ndarray = None
    vp_local: np.ndarray = None
    vp_scan: np.ndarray = None
    cand_ang4: np.ndarray = None
    _dev: dict = field(default_factory=dict)

    @property
    def n_vp(self):
        return int(self.scan_off[-1])

    def gid(self, scan_idx, local):
        return int(self.scan_off[scan_idx]) + int(local)

    def long_id(self, g):
        s = int(self.vp_scan[g])
        return f"{self.scans[s]}_{self.vp_names[s][int(self.vp_local[g])]}"

    # ---- derived tables ---------------------------------------------------
    def build_routes(self):
        """All-pairs shortest paths per scan (common_env.py:164-181 does this with networkx).
        Dijkstra from every source with predecessor tracking; distance accumulates from the
        source outward (dist[v] = dist[u] + w), the same association order as networkx."""
        import heapq
        n_vp = self.n_vp
        self.vp_scan = np.zeros(n_vp, np.int32)
        self.vp_local = np.zeros(n_vp, np.int32)
        self.sq_off = np.zeros(n_vp, np.int64)
        sizes = np.diff(self.scan_off)
        total = int((sizes.astype(np.int64) ** 2).sum())
        self.next_hop = np.full(total, -1, np.int32)
        self.dist = np.zeros(total, np.float32)
        base = 0
        for s, n in enumerate(sizes):
            n = int(n)
            o = int(self.scan_off[s])
            self.vp_scan[o:o + n] = s
            self.vp_local[o:o + n] = np.arange(n)
            self.sq_off[o:o + n] = base + np.arange(n, dtype=np.int64) * n
            adj = [[] for _ in range(n)]
            for (u, v), w in self.edge_len[s].items():
                adj[u].append((v, w)), adj[v].append((u, w))
            for src in range(n):
                d = [math.inf] * n
                first = [-1] * n                     # first hop from src towards each node
                d[src] = 0.0
                first[src] = src
                pq = [(0.0, src)]
                while pq:
                    du, u = heapq.heappop(pq)
                    if du > d[u]:
                        continue
                    for v, w in adj[u]:
                        nd = du + w
                        if nd < d[v]:
                            d[v] = nd
                            first[v] = v if u == src else first[u]
                            heapq.heappush(pq, (nd, v))
                row = base + src * n
                self.dist[row:row + n] = np.asarray(d, np.float64).astype(np.float32)
                self.next_hop[row:row + n] = np.asarray(first, np.int32) + o
            base += n * n
        return self

    def build_cand_angles(self):
        """make_candidate's angle feature for every (viewpoint, neighbour, base heading):
        make_angle_feat(normalized_heading - (viewId % 12)*pi/6, loc_elevation)
        (common_env.py:283-289), evaluated with math.sin/cos in float64 then cast to fp32
        exactly as the reference does — the gather is bit-exact by construction."""
        n_vp = self.n_vp
        out = np.zeros((n_vp, CMAX, 12, 4), np.float32)
        for g in range(n_vp):
            for j in range(int(self.n_cand[g])):
                nh, el = float(self.cand_nheading[g, j]), float(self.cand_elev[g, j])
                se, ce = math.sin(el), math.cos(el)
                for k in range(12):
                    lh = nh - k * ANGLE_INC
                    out[g, j, k] = (math.sin(lh), math.cos(lh), se, ce)
        self.cand_ang4 = out
        return self

    # ---- device face ------------------------------------------------------
    def device_tables(self, device):
        key = str(device)
        if key not in self._dev:
            t = lambda a, dt=None: torch.as_tensor(a, dtype=dt).to(device)
            self._dev[key] = dict(
                table=self.table.to(device),
                cand_vp=t(self.cand_vp, torch.int32), cand_view=t(self.cand_view, torch.int32),
                cand_ang4=t(self.cand_ang4, torch.float32), n_cand=t(self.n_cand, torch.int32),
                next_hop=t(self.next_hop, torch.int32), dist=t(self.dist, torch.float32),
                sq_off=t(self.sq_off, torch.int64), vp_local=t(self.vp_local, torch.int32),
                loc4=t(static_loc4(), torch.float32), pose4=t(pose4(), torch.float32),
            )
        return self._dev[key]

    # ---- host queries (shared by both env faces' tests) -------------------
    def hop(self, cur_g, goal_g):
        return int(self.next_hop[int(self.sq_off[cur_g]) + int(self.vp_local[goal_g])])

    def distance(self, cur_g, goal_g):
        return self.dist[int(self.sq_off[cur_g]) + int(self.vp_local[goal_g])]

    def teacher_action(self, cur_g, goal_g):
        """_teacher_action (base.py:159-178) on indices: slot of the neighbour that is the next
        hop, or n_cand (STOP) at the goal."""
        if cur_g == goal_g:
            return int(self.n_cand[cur_g])
        nh = self.hop(cur_g, goal_g)
        row = self.cand_vp[cur_g, :int(self.n_cand[cur_g])]
        return int(np.nonzero(row == nh)[0][0])


def make_world(n_scans=4, sizes=None, seed=2020, mean_degree=4.0, device="cpu",
               with_table=True, table_seed=None):
    """Random connected scans.  Degree 1..CMAX, edge length U(1, 3.5) m, per directed edge a
    view index U{0..35}, loc_elevation = that view's row elevation, normalized_heading =
    view heading + U(-pi/12, pi/12).  Candidate order = ascending (absViewIndex, neighbour):
    the order a 0..35 view sweep first meets them (common_env.py:235-281)."""
    rng = _pyrandom.Random(seed)
    if sizes is None:
        sizes = [rng.randint(20, 40) for _ in range(n_scans)]
    n_scans = len(sizes)
    scans = [f"scan{idx:03d}" for idx in range(n_scans)]
    vp_names = [[f"vp{idx:03d}x{j:04d}" for j in range(n)] for idx, n in enumerate(sizes)]
    scan_off = np.concatenate(([0], np.cumsum(sizes))).astype(np.int64)
    n_vp = int(scan_off[-1])
    cand_vp = np.full((n_vp, CMAX), -1, np.int32)
    cand_view = np.zeros((n_vp, CMAX), np.int32)
    cand_nh = np.zeros((n_vp, CMAX), np.float64)
    cand_el = np.zeros((n_vp, CMAX), np.float64)
    cand_rel = np.zeros((n_vp, CMAX), np.float64)
    n_cand = np.zeros(n_vp, np.int32)
    edge_len = []
    for s, n in enumerate(sizes):
        deg = [0] * n
        edges = {}
        order = list(range(n))
        rng.shuffle(order)
        for i in range(1, n):                          # random spanning tree
            for _ in range(64):
                u = order[rng.randrange(i)]
                if deg[u] < CMAX:
                    break
            v = order[i]
            edges[(min(u, v), max(u, v))] = rng.uniform(1.0, 3.5)
            deg[u] += 1
            deg[v] += 1
        extra = max(0, int(round(mean_degree * n / 2.0)) - (n - 1))
        tries = 0
        while extra > 0 and tries < 50 * n:
            tries += 1
            u, v = rng.randrange(n), rng.randrange(n)
            if u == v or (min(u, v), max(u, v)) in edges or deg[u] >= CMAX or deg[v] >= CMAX:
                continue
            edges[(min(u, v), max(u, v))] = rng.uniform(1.0, 3.5)
            deg[u] += 1
            deg[v] += 1
            extra -= 1
        edge_len.append(edges)
        nbrs = [[] for _ in range(n)]
        for (u, v) in edges:
            nbrs[u].append(v), nbrs[v].append(u)
        o = int(scan_off[s])
        for u in range(n):
            lst = []
            for v in nbrs[u]:
                view = rng.randrange(N_VIEWS)
                rel = rng.uniform(-math.pi / 12, math.pi / 12)
                lst.append((view, v, view_heading(view) + rel, view_elevation(view), rel))
            lst.sort(key=lambda e: (e[0], e[1]))
            n_cand[o + u] = len(lst)
            for j, (view, v, nh, el, rel) in enumerate(lst):
                cand_vp[o + u, j] = o + v
                cand_view[o + u, j] = view
                cand_nh[o + u, j] = nh
                cand_el[o + u, j] = el
                cand_rel[o + u, j] = rel
    w = World(scans, vp_names, scan_off, cand_vp, cand_view, cand_nh, cand_el, cand_rel, n_cand, edge_len)
    w.build_routes().build_cand_angles()
    if with_table:
        w.table = make_table(n_vp, seed if table_seed is None else table_seed, device)
    return w


def make_table(n_vp, seed, device="cpu", chunk=512):
    """relu(N(0,1)) * 0.8, rounded once to bf16.  Generated chunk-wise on the
    target device so the 1.56 GB full-size table never exists in fp32."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((n_vp, N_VIEWS, IMG_DIM), dtype=torch.bfloat16, device=device)
    for a in range(0, n_vp, chunk):
        b = min(n_vp, a + chunk)
        x = torch.randn((b - a, N_VIEWS, IMG_DIM), generator=g, device=device)
        out[a:b] = (torch.relu(x) * 0.8).to(torch.bfloat16)
    return out


def full_world_sizes(seed=2020, n_scans=90, total=10567):
    """90 scans of 20..345 viewpoints summing to 10 567 (Matterport3D's count)."""
    rng = _pyrandom.Random(seed)
    sizes = [rng.randint(20, 345) for _ in range(n_scans)]
    while sum(sizes) != total:
        i = rng.randrange(n_scans)
        d = 1 if sum(sizes) < total else -1
        if 20 <= sizes[i] + d <= 345:
            sizes[i] += d
    return sizes


def r2r_length_counts():
    """counts[l] = number of R2R training instructions of encoded length l (0..80): the shipped data/R2R_train.json
    tokenised with the reference's rules (oracle/make_length_hist.py wrote the file; mean 31.3, 0.6 % at 80)."""
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "r2r_train_lengths.json")) as f:
        return json.load(f)["counts"]


def make_items(world, n_items, seed=2020, max_len=80, fixed_len=None, vocab=992,
               min_nodes=4, max_nodes=7, instr_per_path=1, length_counts=None):
    """R2R-shaped episodes: shortest path of 4..7 viewpoints, start heading snapped to 30 degrees,
    token ids U{4..vocab-1} framed by <BOS>=3 ... <EOS>=2 and padded with <PAD>=0
    (Tokenizer.encode_sentence, misc.py:139-157)."""
    rng = _pyrandom.Random(seed * 7919 + 13)
    sizes = np.diff(world.scan_off)
    items = []
    pid = 0
    while len(items) < n_items:
        s = rng.choices(range(len(sizes)), weights=[int(x) for x in sizes])[0]
        n = int(sizes[s])
        src = rng.randrange(n)
        path = None
        for _ in range(30):
            dst = rng.randrange(n)
            p = [src]
            g_src, g_dst = world.gid(s, src), world.gid(s, dst)
            cur = g_src
            while cur != g_dst and len(p) <= max_nodes:
                cur = world.hop(cur, g_dst)
                p.append(cur - int(world.scan_off[s]))
            if cur == g_dst and min_nodes <= len(p) <= max_nodes:
                path = p
                break
        if path is None:
            continue
        heading = rng.randrange(12) * ANGLE_INC
        for j in range(instr_per_path):
            if len(items) >= n_items:
                break
            if fixed_len is not None:
                length = fixed_len
            elif length_counts is not None:          # the real R2R length distribution
                length = min(max_len, max(3, rng.choices(range(len(length_counts)), weights=length_counts)[0]))
            else:
                length = min(max_len, max(5, int(rng.gauss(31, 12))))
            enc = np.zeros(max_len, np.int64)
            enc[0] = 3
            for k in range(1, length - 1):
                enc[k] = rng.randrange(4, vocab)
            enc[length - 1] = 2
            items.append({
                "scan": world.scans[s], "scan_idx": s, "path_id": pid, "instr_id": f"{pid}_{j}",
                "path": [world.vp_names[s][u] for u in path],
                "path_g": [world.gid(s, u) for u in path],
                "heading": heading, "instructions": f"synthetic {pid}_{j}",
                "instr_encoding": enc, "instr_length": int(length),
                "distance": float(world.distance(world.gid(s, path[0]), world.gid(s, path[-1]))),
            })
        pid += 1
    return items
