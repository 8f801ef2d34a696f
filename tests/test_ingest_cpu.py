"""Real-data ingestion (environ/ingest.py) round trip on CPU: a world written in the reference's on-disk formats
(connectivity json, feature TSV, candidate cache, R2R json + vocab) loads back into identical index tables."""
import json
import os

import numpy as np
import torch

import clvln_b200  # noqa: F401
from clvln_b200.environ import ingest, make_world
from clvln_b200.environ.batch import R2RBatch


def _write_world(world, root, rng):
    """Connectivity files with 3-D poses (edge length = Euclidean distance, as load_nav_graphs computes it)."""
    os.makedirs(os.path.join(root, "connectivity"), exist_ok=True)
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
        # an excluded panorama that claims links: must not become a node
        pose = [0.0] * 16
        data.append({"image_id": "excluded" + scan, "pose": pose, "included": False, "unobstructed": [True] * n + [False]})
        for d in data[:-1]:
            d["unobstructed"].append(True)
        with open(os.path.join(root, "connectivity", f"{scan}_connectivity.json"), "w") as f:
            json.dump(data, f)


def test_world_round_trip(tmp_path):
    rng = np.random.RandomState(0)
    w0 = make_world(n_scans=2, seed=5)
    root = str(tmp_path)
    _write_world(w0, root, rng)
    feats = {w0.long_id(g): w0.table[g].float().numpy() for g in range(w0.n_vp)}
    ingest.write_feature_tsv(os.path.join(root, "feat.tsv"), feats)
    with open(os.path.join(root, "cands.json"), "w") as f:
        json.dump(ingest.dump_candidates(w0), f)
    w1 = ingest.world_from_files(os.path.join(root, "connectivity"), w0.scans, os.path.join(root, "cands.json"),
                                 os.path.join(root, "feat.tsv"))
    assert w1.n_vp == w0.n_vp
    # viewpoint order may differ (file order of first appearance as a graph node): compare through names
    g1 = {w1.long_id(g): g for g in range(w1.n_vp)}
    perm = np.array([g1[w0.long_id(g)] for g in range(w0.n_vp)])
    assert torch.equal(w1.table[perm].view(torch.int16), w0.table.view(torch.int16))          # bf16 table bit-exact
    assert np.array_equal(w1.n_cand[perm], w0.n_cand)
    assert np.array_equal(w1.cand_view[perm], w0.cand_view)
    assert np.array_equal(w1.cand_ang4[perm], w0.cand_ang4)                                   # float64 sin/cos -> fp32
    for g in range(w0.n_vp):
        k = int(w0.n_cand[g])
        assert np.array_equal(w1.cand_vp[perm[g], :k], perm[w0.cand_vp[g, :k]])
    # distances follow the poses written to disk; next hops are consistent with them
    for a in range(0, w1.n_vp, 7):
        for b in range(0, w1.n_vp, 5):
            if w1.vp_scan[a] != w1.vp_scan[b] or a == b:
                continue
            nh = w1.hop(a, b)
            assert nh in set(w1.cand_vp[a, :int(w1.n_cand[a])].tolist())
            s = int(w1.vp_scan[a])
            o = int(w1.scan_off[s])
            e = w1.edge_len[s][(min(a - o, nh - o), max(a - o, nh - o))]
            assert abs(float(w1.distance(a, b)) - (e + float(w1.distance(nh, b)))) < 1e-4


def test_items_and_tokenizer(tmp_path):
    w = make_world(n_scans=2, seed=5)
    vocab = ["<PAD>", "<UNK>", "<EOS>", "<BOS>", "walk", "to", "the", "door", ",", ".", "stop"]
    with open(tmp_path / "vocab.txt", "w") as f:
        f.write("\n".join(vocab) + "\n")
    tok = ingest.Tokenizer.from_file(str(tmp_path / "vocab.txt"), 12)
    assert tok.split_sentence("Walk to the door, stop!?  ") == ["walk", "to", "the", "door", ",", "stop", "!", "?"]
    enc, n = tok.encode_sentence("Walk to the fridge. Stop")
    assert enc.tolist() == [3, 4, 5, 6, 1, 9, 10, 2, 0, 0, 0, 0] and n == 8
    enc, n = tok.encode_sentence("walk " * 30)
    assert n == 12 and enc[-1] == 2 and len(enc) == 12
    assert tok.encode_sentence("   ") is None
    s = 1
    path = w.vp_names[s][:3]
    data = [{"distance": 5.0, "scan": w.scans[s], "path_id": 77, "path": path, "heading": 3.751,
             "instructions": ["Walk to the door.", "Stop , walk to the door"]},
            {"distance": 1.0, "scan": "unknown_scan", "path_id": 78, "path": path, "heading": 0.0, "instructions": ["walk"]}]
    with open(tmp_path / "R2R_train.json", "w") as f:
        json.dump(data, f)
    items = ingest.items_from_r2r_json(str(tmp_path / "R2R_train.json"), w, tok)
    assert [it["instr_id"] for it in items] == ["77_0", "77_1"]
    assert items[0]["path_g"] == [w.gid(s, j) for j in range(3)] and items[0]["instr_length"] == 7
    assert ingest.heading_to_start_view(3.751) == 12 + 7 and ingest.heading_to_start_view(6.2) == 12
    env = R2RBatch(w, items, batch_size=2)                       # the converted items drive the index-table env
    assert len(env.data) == 2
