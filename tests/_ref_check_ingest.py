"""environ/ingest.py vs the REAL reference loaders (container only): feature TSV reader, connectivity graphs +
networkx shortest paths (ties included), tokenizer on the shipped R2R instructions."""
import base64
import json
import os
import sys
import tempfile

import networkx as nx
import numpy as np

sys.path.insert(0, "/root/repo")
import clvln_b200  # noqa: E402,F401
from clvln_b200.environ import ingest  # noqa: E402
from oracle import ref_loader  # noqa: E402

ref_loader.load_ref_agents()                       # installs the stubs misc.py's imports need
import importlib.util  # noqa: E402
# a pristine copy of the reference's misc.py: other tests monkeypatch src.utils.misc.load_nav_graphs in this process
_spec = importlib.util.spec_from_file_location("ref_misc_pristine", os.path.join(ref_loader.REF_TASK, "src", "utils", "misc.py"))
M = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(M)

if not hasattr(base64, "decodestring"):          # removed in Python 3.9; read_in (misc.py:274) still calls it
    base64.decodestring = base64.decodebytes

rng = np.random.RandomState(3)
tmp = tempfile.mkdtemp()
os.makedirs(os.path.join(tmp, "connectivity"))

# ---- a scan on a 5x4 grid (unit spacing: many equal-length shortest paths), one excluded node, some missing links ----
scan = "gridscan"
W_, H_ = 5, 4
ids = ["%032x" % (1000 + k) for k in range(W_ * H_)]
data = []
for k in range(W_ * H_):
    x, y = k % W_, k // W_
    pose = [0.0] * 16
    pose[3], pose[7], pose[11] = float(x), float(y), 1.5
    unob = [False] * (W_ * H_)
    for (dx, dy) in ((1, 0), (-1, 0), (0, 1), (0, -1)):
        xx, yy = x + dx, y + dy
        if 0 <= xx < W_ and 0 <= yy < H_ and not ({k, yy * W_ + xx} == {6, 7}):
            unob[yy * W_ + xx] = True
    data.append({"image_id": ids[k], "pose": pose, "included": k != 13, "unobstructed": unob})
order = list(rng.permutation(W_ * H_))            # file order unrelated to the grid order
remap = {old: new for new, old in enumerate(order)}
data2 = []
for old in order:
    it = dict(data[old])
    un = [False] * (W_ * H_)
    for j, c in enumerate(data[old]["unobstructed"]):
        if c:
            un[remap[j]] = True
    it["unobstructed"] = un
    data2.append(it)
with open(os.path.join(tmp, "connectivity", f"{scan}_connectivity.json"), "w") as f:
    json.dump(data2, f)

cwd = os.getcwd()
os.chdir(tmp)
try:
    G = M.load_nav_graphs([scan])[scan]
finally:
    os.chdir(cwd)
ref_paths = dict(nx.all_pairs_dijkstra_path(G))                    # common_env.py:170-175
ref_dist = dict(nx.all_pairs_dijkstra_path_length(G))

names, edges = ingest.load_connectivity(os.path.join(tmp, "connectivity"), scan)
assert set(names) == set(G.nodes) and len(edges) == G.number_of_edges()
cache = {f"{scan}_{vp}": [] for vp in names}
world = ingest.world_from_files(os.path.join(tmp, "connectivity"), [scan], cache)
ties = 0
for a, va in enumerate(names):
    for b, vb in enumerate(names):
        p = ref_paths[va][vb]
        want = p[1] if len(p) > 1 else va
        got = names[world.hop(a, b)]
        assert got == want, (va, vb, got, want)
        assert world.distance(a, b) == np.float32(ref_dist[va][vb])
        ties += len(list(nx.all_shortest_paths(G, va, vb, weight="weight"))) > 1
assert ties > 50, ties
print(f"connectivity: {len(names)} nodes, {len(edges)} edges, next hop identical on all pairs ({ties} pairs with ties)")

# ---- feature TSV ----
feats = {f"{scan}_{vp}": rng.randn(36, 2048).astype(np.float32) for vp in names[:5]}
tsv = os.path.join(tmp, "feat.tsv")
ingest.write_feature_tsv(tsv, feats)
ref_feats = M.ImageFeatures.read_in(tsv)
mine = ingest.read_feature_tsv(tsv)
assert set(ref_feats) == set(mine) == set(feats)
for k in feats:
    assert np.array_equal(ref_feats[k], mine[k]) and np.array_equal(mine[k], feats[k])
print("feature TSV: reader identical to ImageFeatures.read_in")

# ---- tokenizer on the shipped R2R instructions ----
data_dir = os.path.join(ref_loader.REF_TASK, "data")
vocab = M.read_vocab(os.path.join(data_dir, "train_vocab.txt"))
rt = M.Tokenizer(vocab=vocab, encoding_length=80)
mt = ingest.Tokenizer.from_file(os.path.join(data_dir, "train_vocab.txt"), 80)
assert mt.vocab_size() == rt.vocab_size() == 992
n = 0
for item in json.load(open(os.path.join(data_dir, "R2R_val_unseen.json")))[:400]:
    for instr in item["instructions"]:
        a, b = rt.encode_sentence(instr), mt.encode_sentence(instr)
        assert (a is None) == (b is None)
        if a is not None:
            assert np.array_equal(a[0], b[0]) and a[1] == b[1]
            n += 1
long = "walk , " * 100
a, b = rt.encode_sentence(long), mt.encode_sentence(long)
assert np.array_equal(a[0], b[0]) and a[1] == b[1] == 80 and b[0][-1] == mt.word_to_index["<EOS>"]
print(f"tokenizer: {n} instructions encode identically (incl. unknown words and truncation)")

# ---- candidate cache: the REAL R2RBatch.make_candidate (over the table-driven fake simulator) -> dump -> tables ----
import random  # noqa: E402
from clvln_b200.environ import make_items, make_world  # noqa: E402
from oracle import ref_harness as H  # noqa: E402

w = make_world(n_scans=2, seed=9)
items = make_items(w, 24, seed=9)
rsrc = H.install(w, {"train": items})
import src.environ as renv  # noqa: E402
random.seed(2020)
ref_env = renv.R2RBatch(H.feature_store(w), batch_size=4, splits=["train"], tokenizer=H.StubTokenizer(items))
dump = os.path.join(tmp, "cands.json")
ingest.dump_candidates_with_reference(ref_env, {s: w.vp_names[i] for i, s in enumerate(w.scans)}, dump)
w2 = make_world(n_scans=2, seed=9, with_table=False)
for name in ("cand_vp", "cand_view", "cand_nheading", "cand_elev", "n_cand"):
    getattr(w2, name)[...] = 0                         # wipe, then refill from the reference's cache
ingest.candidates_to_tables(w2, json.load(open(dump)))
w2.build_cand_angles()
assert np.array_equal(w2.n_cand, w.n_cand) and np.array_equal(w2.cand_view, w.cand_view)
for g in range(w.n_vp):
    k = int(w.n_cand[g])
    assert np.array_equal(w2.cand_vp[g, :k], w.cand_vp[g, :k])
    assert np.allclose(w2.cand_elev[g, :k], w.cand_elev[g, :k], rtol=0, atol=1e-12)
    assert np.allclose(w2.cand_nheading[g, :k], w.cand_nheading[g, :k], rtol=0, atol=1e-9)
assert np.abs(w2.cand_ang4 - w.cand_ang4).max() < 1e-6
print(f"candidate cache: reference make_candidate over {w.n_vp} viewpoints -> dump -> tables identical")
