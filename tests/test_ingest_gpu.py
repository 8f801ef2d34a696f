"""Real-data ingestion on the GPU: the reference's on-disk files -> HBM tables -> the gather kernels.

A world is written in the reference's formats (feature TSV with base64 fp32 [36,2048] rows as ImageFeatures.read_in
parses them, misc.py:254-279; connectivity json; the candidate cache), read back through environ/ingest.py (and through
the `src`-compatible facade), uploaded, and `vln_gather_pano` / `vln_gather_cand` must return — bit for bit — what the
reference's observe() would assemble from the decoded TSV: the features (rounded once to bf16, the table's storage type)
concatenated with the float64-computed angle embeddings."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_tsv_to_hbm_table_to_gather_is_bit_exact(tmp_path, monkeypatch):
    import clvln_b200  # noqa: F401
    import clvln_b200.compat as compat
    from clvln_b200 import ops
    from clvln_b200.environ import ingest, make_world
    from clvln_b200.environ.world import static_loc4
    dev = torch.device("cuda:0")
    w0 = make_world(n_scans=2, seed=9)
    rs = np.random.RandomState(0)
    # features that are NOT bf16-exact: the conversion's round-to-nearest-even is part of what is checked
    feats = {w0.long_id(g): np.maximum(rs.randn(36, 2048), 0).astype(np.float32) * 0.8 for g in range(w0.n_vp)}
    ingest.write_connectivity(w0, str(tmp_path / "connectivity"))
    os.makedirs(tmp_path / "img_features")
    tsv = str(tmp_path / "img_features" / "ResNet-152-imagenet.tsv")
    ingest.write_feature_tsv(tsv, feats)
    with open(tmp_path / "img_features" / "candidates.json", "w") as f:
        json.dump(ingest.dump_candidates(w0), f)
    monkeypatch.chdir(tmp_path)
    # what ImageFeatures.read_in would hold: the decoded rows, bit for bit
    decoded = ingest.read_feature_tsv(tsv)
    assert all(np.array_equal(decoded[k], feats[k]) for k in feats)
    world = compat.utils.ImageFeatures.read_in(tsv).world()           # facade path: TSV + connectivity + candidate cache
    store = ops.FeatureStore.from_world(world, dev)
    vp = torch.arange(world.n_vp, dtype=torch.int32, device=dev)
    view = (vp * 7 % 36).to(torch.int32)
    got = ops.gather_pano(store, vp, view).cpu()
    loc4 = torch.from_numpy(static_loc4())
    for g in range(world.n_vp):
        exp_img = torch.from_numpy(decoded[world.long_id(g)]).to(torch.bfloat16).float()      # RNE, once
        exp = torch.cat((exp_img, loc4[int(view[g])].repeat_interleave(32, dim=1)), 1)
        assert torch.equal(got[g], exp), g
    # candidates: rows table[vp, absViewIndex] + the angle feature of (normalized_heading - base_heading, elevation)
    cand, lens = ops.gather_cand(store, vp, view)
    cand, lens = cand.cpu(), lens.cpu()
    for g in range(0, world.n_vp, 3):
        n = int(world.n_cand[g])
        assert int(lens[g]) == n + 1 and bool((cand[g, n:] == 0).all())
        for j in range(n):
            av = int(world.cand_view[g, j])
            exp_img = torch.from_numpy(decoded[world.long_id(g)][av]).to(torch.bfloat16).float()
            ang = torch.from_numpy(world.cand_ang4[g, j, int(view[g]) % 12]).repeat_interleave(32)
            assert torch.equal(cand[g, j], torch.cat((exp_img, ang)))
