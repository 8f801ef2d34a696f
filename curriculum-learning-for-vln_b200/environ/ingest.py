"""Real-data ingestion: the reference's on-disk formats -> the HBM index tables of `World`.

SURVEY.md §8(f) rank 2.  Nothing here runs on the hot path: it is the one-time conversion a
user does when the Matterport3D features / connectivity graphs / R2R json are available, and it
produces exactly the tables the synthetic `make_world` emits, so everything downstream (FeatureStore,
R2RBatch, the agents) is unchanged.  Formats and the reference code that defines them:

  * image features  img_features/ResNet-152-imagenet.tsv — tab-separated rows
        scanId, viewpointId, image_w, image_h, vfov, features
    with `features` = base64 of float32[36, 2048]           (ImageFeatures.read_in, utils/misc.py:254-279;
    its `base64.decodestring` no longer exists in Python >= 3.9 — `decodebytes` is the same function)
  * connectivity/<scan>_connectivity.json — list of {image_id, pose[16], included, unobstructed[]}
    (load_nav_graphs, utils/misc.py:35-60; shortest paths / distances with networkx exactly as
    R2RBatch._load_nav_graphs does, environ/common_env.py:164-181, so ties break the same way)
  * the candidate cache — R2RBatch.make_candidate (common_env.py:225-297) sweeps the 36 views of the
    Matterport simulator per (scan, viewpoint) and keeps, per navigable neighbour, the view with the
    smallest angular distance.  The simulator (MatterSim, C++, not vendored) is the only place that
    geometry is defined, so it is NOT re-derived here: the converter reads a JSON dump of the reference's
    own `buffered_state_dict`, verbatim (common_env.py:276-281) — {"<scan>_<viewpoint>": [{"scanId",
    "absViewIndex", "nextViewpointId", "normalized_heading", "loc_elevation", "distance", "idx"}, ...]} — which
    `dump_candidates_with_reference` writes by running the reference's make_candidate once over every viewpoint
    (needs MatterSim; run on the user's side).
  * data/R2R_<split>.json + data/train_vocab.txt — episodes and the tokenizer vocabulary
    (load_datasets utils/misc.py:62-69, Tokenizer :91-157, R2RBatch.__init__ common_env.py:121-150).
"""
import base64
import csv
import json
import os
import re
import string
import sys

import numpy as np
import torch

from .world import CMAX, IMG_DIM, N_VIEWS, World, heading_to_view

TSV_FIELDS = ["scanId", "viewpointId", "image_w", "image_h", "vfov", "features"]
IMAGE_W, IMAGE_H, VFOV = 640, 480, 60          # ImageFeatures constants, misc.py:244-250


# ---- image features -----------------------------------------------------------------------------------
def iter_feature_tsv(path, views=N_VIEWS):
    """Yield (long_id, float32[views, 2048]) per TSV row; same checks as ImageFeatures.read_in."""
    csv.field_size_limit(sys.maxsize)
    with open(path, "r") as f:
        for item in csv.DictReader(f, delimiter="\t", fieldnames=TSV_FIELDS):
            assert int(item["image_h"]) == IMAGE_H and int(item["image_w"]) == IMAGE_W and int(item["vfov"]) == VFOV
            feat = np.frombuffer(base64.decodebytes(item["features"].encode("ascii")), dtype=np.float32)
            yield item["scanId"] + "_" + item["viewpointId"], feat.reshape((views, -1))


def read_feature_tsv(path, views=N_VIEWS):
    """The dict ImageFeatures.read_in returns: "<scan>_<viewpoint>" -> float32[36, 2048]."""
    return dict(iter_feature_tsv(path, views))


def write_feature_tsv(path, features):
    """Inverse of read_feature_tsv (fixtures and tests): `features` maps long ids to float32[36, 2048]."""
    with open(path, "w") as f:
        for long_id, feat in features.items():
            scan, vp = long_id.split("_", 1)
            b64 = base64.b64encode(np.ascontiguousarray(feat, dtype=np.float32).tobytes()).decode("ascii")
            f.write("\t".join([scan, vp, str(IMAGE_W), str(IMAGE_H), str(VFOV), b64]) + "\n")


def feature_table(world, features, device="cpu"):
    """bf16 [n_vp, 36, 2048] in the world's viewpoint order (round-to-nearest-even, once): the table
    FeatureStore keeps in HBM.  Viewpoints without a TSV row are an error (the reference would raise a
    KeyError at the first lookup, common_env.py:80)."""
    table = torch.empty((world.n_vp, N_VIEWS, IMG_DIM), dtype=torch.bfloat16, device=device)
    for s, scan in enumerate(world.scans):
        o = int(world.scan_off[s])
        for j, vp in enumerate(world.vp_names[s]):
            feat = features[f"{scan}_{vp}"]
            assert feat.shape == (N_VIEWS, IMG_DIM), feat.shape
            table[o + j] = torch.from_numpy(np.array(feat, dtype=np.float32)).to(device).to(torch.bfloat16)
    return table


# ---- connectivity ------------------------------------------------------------------------------------
def _pose_distance(a, b):
    return ((a["pose"][3] - b["pose"][3]) ** 2 + (a["pose"][7] - b["pose"][7]) ** 2
            + (a["pose"][11] - b["pose"][11]) ** 2) ** 0.5


def load_connectivity(conn_dir, scan):
    """(viewpoint ids in file order restricted to graph nodes, {(u, v): metres} with u < v) of one scan:
    the nodes / weighted edges load_nav_graphs (misc.py:35-60) puts into its networkx graph."""
    with open(os.path.join(conn_dir, f"{scan}_connectivity.json")) as f:
        data = json.load(f)
    names, index, edges = [], {}, {}
    for i, item in enumerate(data):
        if not item["included"]:
            continue
        for j, conn in enumerate(item["unobstructed"]):
            if conn and data[j]["included"]:
                assert data[j]["unobstructed"][i], "Graph should be undirected"
                for it in (item, data[j]):
                    if it["image_id"] not in index:
                        index[it["image_id"]] = len(names)
                        names.append(it["image_id"])
                u, v = index[item["image_id"]], index[data[j]["image_id"]]
                edges[(min(u, v), max(u, v))] = _pose_distance(item, data[j])
    return names, edges


def _routes_networkx(world):
    """next_hop / dist tables through networkx, with the calls of R2RBatch._load_nav_graphs
    (common_env.py:164-181: all_pairs_dijkstra_path / all_pairs_dijkstra_path_length), so equal-length
    paths resolve to the SAME next viewpoint the reference's teacher would pick."""
    import networkx as nx
    n_vp = world.n_vp
    sizes = np.diff(world.scan_off)
    world.vp_scan = np.zeros(n_vp, np.int32)
    world.vp_local = np.zeros(n_vp, np.int32)
    world.sq_off = np.zeros(n_vp, np.int64)
    total = int((sizes.astype(np.int64) ** 2).sum())
    world.next_hop = np.full(total, -1, np.int32)
    world.dist = np.zeros(total, np.float32)
    base = 0
    for s, n in enumerate(sizes):
        n, o = int(n), int(world.scan_off[s])
        world.vp_scan[o:o + n] = s
        world.vp_local[o:o + n] = np.arange(n)
        world.sq_off[o:o + n] = base + np.arange(n, dtype=np.int64) * n
        G = nx.Graph()
        # insertion order = load_nav_graphs': edges in (file order of item, order of its unobstructed list)
        for (u, v), w in world.edge_len[s].items():
            G.add_edge(u, v, weight=w)
        paths = dict(nx.all_pairs_dijkstra_path(G))
        dists = dict(nx.all_pairs_dijkstra_path_length(G))
        for src in range(n):
            row = base + src * n
            for dst in range(n):
                p = paths[src][dst]
                world.next_hop[row + dst] = o + (p[1] if len(p) > 1 else src)
                world.dist[row + dst] = np.float32(dists[src][dst])
        base += n * n
    return world


# ---- candidates ----------------------------------------------------------------------------------------
def candidates_to_tables(world, cache):
    """`cache`: the reference's buffered_state_dict (long id -> candidate list in make_candidate's order,
    common_env.py:258-281) -> cand_vp / cand_view / normalized heading / elevation / n_cand."""
    n_vp = world.n_vp
    world.cand_vp = np.full((n_vp, CMAX), -1, np.int32)
    world.cand_view = np.zeros((n_vp, CMAX), np.int32)
    world.cand_nheading = np.zeros((n_vp, CMAX), np.float64)
    world.cand_elev = np.zeros((n_vp, CMAX), np.float64)
    world.cand_rel = np.zeros((n_vp, CMAX), np.float64)
    world.n_cand = np.zeros(n_vp, np.int32)
    for s, scan in enumerate(world.scans):
        o = int(world.scan_off[s])
        local = {vp: j for j, vp in enumerate(world.vp_names[s])}
        for j, vp in enumerate(world.vp_names[s]):
            cands = cache[f"{scan}_{vp}"]
            assert len(cands) <= CMAX, f"{scan}_{vp}: {len(cands)} candidates (table width {CMAX})"
            world.n_cand[o + j] = len(cands)
            for k, c in enumerate(cands):
                world.cand_vp[o + j, k] = o + local[c["nextViewpointId"]]
                world.cand_view[o + j, k] = int(c["absViewIndex"])
                world.cand_nheading[o + j, k] = float(c["normalized_heading"])
                world.cand_elev[o + j, k] = float(c["loc_elevation"])
    return world


def dump_candidates(world):
    """The cache JSON of a world (what dump_candidates_with_reference would write for the real one)."""
    out = {}
    for g in range(world.n_vp):
        s = int(world.vp_scan[g])
        o = int(world.scan_off[s])
        out[world.long_id(g)] = [
            {"scanId": world.scans[s], "absViewIndex": int(world.cand_view[g, k]),
             "nextViewpointId": world.vp_names[s][int(world.cand_vp[g, k]) - o],
             "normalized_heading": float(world.cand_nheading[g, k]), "loc_elevation": float(world.cand_elev[g, k]),
             "distance": 0.0, "idx": k + 1}
            for k in range(int(world.n_cand[g]))]
    return out


def dump_candidates_with_reference(ref_env, scans_to_viewpoints, path):
    """Run on the user's side, with the reference importable and MatterSim installed: fills the reference
    R2RBatch's `buffered_state_dict` by calling its make_candidate(feature, scan, viewpoint, viewId=0)
    (common_env.py:225-297) for every viewpoint and writes the cache as JSON (features dropped)."""
    def plain(v):                       # numpy scalars -> json
        return v.item() if hasattr(v, "item") else v
    out = {}
    for scan, vps in scans_to_viewpoints.items():
        for vp in vps:
            long_id = f"{scan}_{vp}"
            ref_env.make_candidate(ref_env.env.features[long_id], scan, vp, 0)
            out[long_id] = [{k: plain(v) for k, v in c.items()} for c in ref_env.buffered_state_dict[long_id]]
    with open(path, "w") as f:
        json.dump(out, f)
    return out


# ---- world ---------------------------------------------------------------------------------------------
def world_from_files(conn_dir, scans, candidates_json, feature_tsv=None, device="cpu"):
    """World (all index tables + the bf16 feature table) from the reference's files."""
    scans = list(scans)
    names, edges = [], []
    for scan in scans:
        n, e = load_connectivity(conn_dir, scan)
        names.append(n), edges.append(e)
    scan_off = np.concatenate(([0], np.cumsum([len(n) for n in names]))).astype(np.int64)
    z = np.zeros((int(scan_off[-1]), CMAX))
    world = World(scans, names, scan_off, z.astype(np.int32), z.astype(np.int32), z.copy(), z.copy(), z.copy(),
                  np.zeros(int(scan_off[-1]), np.int32), edges)
    _routes_networkx(world)
    cache = candidates_json
    if isinstance(cache, (str, os.PathLike)):
        with open(cache) as f:
            cache = json.load(f)
    candidates_to_tables(world, cache)
    world.build_cand_angles()
    if feature_tsv is not None:
        world.table = feature_table(world, read_feature_tsv(feature_tsv), device)
    return world


# ---- instructions --------------------------------------------------------------------------------------
class Tokenizer:
    """Tokenizer of utils/misc.py:91-157: split on non-alphanumerics, lower-case, punctuation runs broken
    into characters (except runs of full stops), <BOS> ... <EOS>, <PAD> to `encoding_length`, truncation
    ends with <EOS>; unknown words map to <UNK>."""
    SPLIT = re.compile(r"(\W+)")

    def __init__(self, vocab, encoding_length=80):
        self.vocab = list(vocab)
        self.encoding_length = encoding_length
        self.word_to_index = {w: i for i, w in enumerate(self.vocab)}
        self.unk = self.word_to_index["<UNK>"]

    @classmethod
    def from_file(cls, path, encoding_length=80):
        with open(path) as f:
            return cls([w.strip() for w in f.readlines()], encoding_length)      # read_vocab, misc.py:215-218

    def vocab_size(self):
        return len(self.word_to_index)

    @classmethod
    def split_sentence(cls, sentence):
        toks = []
        for word in [s.strip().lower() for s in cls.SPLIT.split(sentence.strip()) if len(s.strip()) > 0]:
            if all(c in string.punctuation for c in word) and not all(c in "." for c in word):
                toks += list(word)
            else:
                toks.append(word)
        return toks

    def encode_sentence(self, sentence, max_length=None):
        max_length = self.encoding_length if max_length is None else max_length
        enc = [self.word_to_index["<BOS>"]]
        enc += [self.word_to_index.get(w, self.unk) for w in self.split_sentence(sentence)]
        enc.append(self.word_to_index["<EOS>"])
        if len(enc) <= 2:
            return None
        if len(enc) < max_length:
            length = len(enc)
            enc += [self.word_to_index["<PAD>"]] * (max_length - len(enc))
        else:
            length = max_length
            enc[max_length - 1] = self.word_to_index["<EOS>"]
        return np.array(enc[:max_length]), length


def items_from_r2r_json(path, world, tokenizer):
    """Episode dicts (the fields R2RBatch consumes) from data/<name>_<split>.json: one item per instruction,
    instr_id = "<path_id>_<j>", scans without features skipped (common_env.py:129-143)."""
    with open(path) as f:
        data = json.load(f)
    scan_idx = {s: i for i, s in enumerate(world.scans)}
    items = []
    for item in data:
        if item["scan"] not in scan_idx:
            continue
        s = scan_idx[item["scan"]]
        local = {vp: j for j, vp in enumerate(world.vp_names[s])}
        path_g = [world.gid(s, local[vp]) for vp in item["path"]]
        for j, instr in enumerate(item["instructions"]):
            enc = tokenizer.encode_sentence(instr)
            if enc is None:
                continue
            new = dict(item)
            new.update({"scan_idx": s, "instr_id": "%s_%d" % (item["path_id"], j), "instructions": instr,
                        "instr_encoding": np.asarray(enc[0], np.int64), "instr_length": int(enc[1]), "path_g": path_g})
            items.append(new)
    return items


def heading_to_start_view(heading):
    """viewIndex after newEpisode(scan, vp, heading, 0): the simulator snaps the heading to the 30-degree grid
    on elevation row 1; what R2RBatch does with each item's `heading`."""
    return heading_to_view(heading)


# ---- the inverse direction: a World + episodes written in the reference's on-disk formats ------------------------
def write_connectivity(world, conn_dir, seed=0):
    """connectivity/<scan>_connectivity.json files for a World (fixtures / tests): every viewpoint gets a random 3-D
    position, `unobstructed` is the World's adjacency; load_nav_graphs recomputes edge lengths from the poses, so a
    world loaded back from these files is self-consistent (distances follow the written poses)."""
    rng = np.random.RandomState(seed)
    os.makedirs(conn_dir, exist_ok=True)
    for s, scan in enumerate(world.scans):
        n = len(world.vp_names[s])
        pos = rng.uniform(0, 20, size=(n, 3))
        nbr = [set() for _ in range(n)]
        for (u, v) in world.edge_len[s]:
            nbr[u].add(v), nbr[v].add(u)
        data = []
        for u in range(n):
            pose = [0.0] * 16
            pose[3], pose[7], pose[11] = (float(x) for x in pos[u])
            data.append({"image_id": world.vp_names[s][u], "pose": pose, "included": True,
                         "unobstructed": [v in nbr[u] for v in range(n)]})
        with open(os.path.join(conn_dir, f"{scan}_connectivity.json"), "w") as f:
            json.dump(data, f)


def synthetic_vocab(vocab=992):
    """<PAD> <UNK> <EOS> <BOS> w4 ... w991: the vocabulary under which make_items' token ids are ordinary words."""
    return ["<PAD>", "<UNK>", "<EOS>", "<BOS>"] + ["w%d" % i for i in range(4, vocab)]


def write_reference_dataset(world, splits, root, dataset="R2R", data_dir="tasks/R2R-judy/data", vocab=992, seed=0):
    """Everything the reference's main.py reads, for a (synthetic) World and {split name: episode items}: under `root`
    connectivity/*.json, img_features/ResNet-152-imagenet.tsv (+ candidates.json next to it: the candidate cache the
    Matterport simulator would produce), <data_dir>/<dataset>_<split>.json with the instructions as text, and the two
    vocabulary files.  Returns the paths main.py's config needs."""
    write_connectivity(world, os.path.join(root, "connectivity"), seed)
    feat_dir = os.path.join(root, "img_features")
    os.makedirs(feat_dir, exist_ok=True)
    tsv = os.path.join(feat_dir, "ResNet-152-imagenet.tsv")
    write_feature_tsv(tsv, {world.long_id(g): world.table[g].float().cpu().numpy() for g in range(world.n_vp)})
    with open(os.path.join(feat_dir, "candidates.json"), "w") as f:
        json.dump(dump_candidates(world), f)
    ddir = os.path.join(root, data_dir)
    os.makedirs(ddir, exist_ok=True)
    for split, items in splits.items():
        by_path = {}
        for it in items:
            text = " ".join("w%d" % int(t) for t in it["instr_encoding"][1:it["instr_length"] - 1])
            d = by_path.setdefault(it["path_id"], {"scan": it["scan"], "path_id": it["path_id"], "path": list(it["path"]),
                                                   "heading": float(it["heading"]), "distance": float(it["distance"]),
                                                   "instructions": []})
            d["instructions"].append(text)
        with open(os.path.join(ddir, "%s_%s.json" % (dataset, split)), "w") as f:
            json.dump(list(by_path.values()), f)
    words = synthetic_vocab(vocab)
    out = {"tsv": tsv, "data_dir": ddir}
    for name in ("train_vocab.txt", "trainval_vocab.txt"):
        out[name] = os.path.join(ddir if dataset == "R2R" else os.path.dirname(ddir), name)
        with open(out[name], "w") as f:
            f.write("\n".join(words) + "\n")
    return out
