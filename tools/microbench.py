"""Config-5 microbench: fused gather + panorama attention (and candidate logits) HBM sweep.

Random viewpoints over the full-size table (1.56 GB >> 126 MB L2).  Each measurement replays a
CUDA graph of N_SETS launches over different random index sets (so launches read HBM, not L2, and
the host launch path is out of the timed region), CUDA events around the replays.
One JSON line per (B, kernel variant, mode, drop): achieved algorithmic GB/s = B*147456 / time per launch."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clvln_b200  # noqa: E402,F401
from clvln_b200 import ops  # noqa: E402
from clvln_b200.environ import world as W  # noqa: E402

N_SETS = 16


def time_graph(launch, reps=10):
    """launch(k) enqueues the k-th variant; returns seconds per launch."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for k in range(2):
            launch(k)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for k in range(N_SETS):
            launch(k)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / (reps * N_SETS)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-vp", type=int, default=10567)
    ap.add_argument("--batches", type=int, nargs="*", default=[16, 64, 128, 256, 512, 1024, 2048])
    ap.add_argument("--splits", type=int, nargs="*", default=[1, 2, 4])
    ap.add_argument("--drops", type=float, nargs="*", default=[0.0, 0.3])
    ap.add_argument("--modes", type=int, nargs="*", default=[0, 1], help="0 forward, 1 backward")
    ap.add_argument("--no-cand", action="store_true")
    ap.add_argument("--cands", type=int, nargs="*", default=[1, 2, 4, 8, 12, 15],
                    help="candidate counts of the cand_logits sweep (BASELINE config 5: candidates 1-16 incl. the END slot)")
    ap.add_argument("--cand-batches", type=int, nargs="*", default=[64, 256, 1024])
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    peaks = {}
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    peak = peaks.get("hbm_gbs", 6650.0)
    n_vp = args.n_vp
    g = torch.Generator(device=dev).manual_seed(1)
    tables = dict(
        table=W.make_table(n_vp, 1, dev),
        cand_vp=torch.randint(0, n_vp, (n_vp, 15), device=dev, dtype=torch.int32, generator=g),
        cand_view=torch.randint(0, 36, (n_vp, 15), device=dev, dtype=torch.int32, generator=g),
        cand_ang4=torch.rand((n_vp, 15, 12, 4), device=dev, generator=g),
        n_cand=torch.randint(1, 16, (n_vp,), device=dev, dtype=torch.int32, generator=g),
        next_hop=torch.zeros(1, device=dev, dtype=torch.int32), dist=torch.zeros(1, device=dev),
        sq_off=torch.zeros(n_vp, device=dev, dtype=torch.int64), vp_local=torch.zeros(n_vp, device=dev, dtype=torch.int32),
        loc4=torch.from_numpy(W.static_loc4()).to(dev), pose4=torch.from_numpy(W.pose4()).to(dev))
    store = ops.FeatureStore(tables, dev)
    rng = ops.Rng(1, dev)
    for B in args.batches:
        vps = [torch.randint(0, n_vp, (B,), device=dev, dtype=torch.int32, generator=g) for _ in range(N_SETS)]
        view = torch.randint(0, 36, (B,), device=dev, dtype=torch.int32, generator=g)
        q = torch.randn(B, 2176, device=dev) * 0.05
        attn = torch.empty(B, 36, device=dev)
        fwd = torch.randn(B, 2176, device=dev)
        for drop in args.drops:
            bits = None
            if drop > 0:      # keep-bits as the rollout pre-generates them (vln_feature_mask_bits), one set per launch
                bits = torch.empty((N_SETS, B * 36, 256), dtype=torch.uint8, device=dev)
                ops._call("vln_feature_mask_bits", ops._ptr(bits), B * 36, N_SETS, drop, rng.ptr, 1, 7, ops._stream())
            for split in args.splits:      # kernel variant: 1 automatic, 2 cluster (low latency), 4 streaming
                for mode in args.modes:
                    out = torch.empty(B, 2176, device=dev)

                    def run(k):
                        ops._call("vln_pano_attn_ld", store.handle, ops._ptr(vps[k]), ops._ptr(view), ops._ptr(store.loc4),
                                  ops._ptr(q), 2176, ops._ptr(attn), None, 2176, ops._ptr(out), 2176, B, mode, drop, rng.ptr, 0,
                                  ops._ptr(bits[k]) if bits is not None else None, split, ops._stream())
                    if mode == 1:
                        attn.copy_(torch.softmax(torch.randn(B, 36, device=dev), 1))
                    t = time_graph(run)
                    gbs = B * 147456 / t / 1e9
                    print(json.dumps(dict(kernel="pano_attn", mode="fwd" if mode == 0 else "bwd", B=B, variant=split,
                                          drop=drop, us=round(t * 1e6, 2), algo_GBs=round(gbs, 1),
                                          frac_of_measured_peak=round(gbs / peak, 3))), flush=True)
            del bits
    if args.no_cand:
        return
    # ---- candidate logits, fwd + bwd, swept over the number of navigable candidates per episode (config 5) ----
    # every viewpoint of a launch has exactly n_cand candidates: algorithmic bytes = B * n_cand * 2048 * 2 each way
    nc_tab = store.n_cand
    for B in args.cand_batches:
        vps = [torch.randint(0, n_vp, (B,), device=dev, dtype=torch.int32, generator=g) for _ in range(N_SETS)]
        view = torch.randint(0, 36, (B,), device=dev, dtype=torch.int32, generator=g)
        tgt = torch.randn(B, 2176, device=dev) * 0.05
        logits = torch.empty(B, 16, device=dev)
        dlogits = torch.randn(B, 16, device=dev)
        d_tgt = torch.empty(B, 2176, device=dev)
        for nc in args.cands:
            nc_tab.fill_(nc)
            for drop in (0.0, 0.3):
                def run_f(k):
                    ops._call("vln_cand_logits_fwd", store.handle, ops._ptr(vps[k]), ops._ptr(view), ops._ptr(store.cand_view),
                              ops._ptr(store.cand_ang4), ops._ptr(nc_tab), ops._ptr(tgt), None, ops._ptr(logits), B, drop,
                              rng.ptr, 3 + k, ops._stream())

                def run_b(k):
                    ops._call("vln_cand_logits_bwd", store.handle, ops._ptr(vps[k]), ops._ptr(view), ops._ptr(store.cand_view),
                              ops._ptr(store.cand_ang4), ops._ptr(nc_tab), ops._ptr(dlogits), ops._ptr(d_tgt), None, B, drop,
                              rng.ptr, 3 + k, ops._stream())
                for name, fn in (("cand_logits_fwd", run_f), ("cand_logits_bwd", run_b)):
                    t = time_graph(fn)
                    gbs = B * nc * 4096 / t / 1e9
                    print(json.dumps(dict(kernel=name, B=B, n_cand=nc, drop=drop, us=round(t * 1e6, 2), algo_GBs=round(gbs, 1),
                                          frac_of_measured_peak=round(gbs / peak, 4))), flush=True)


if __name__ == "__main__":
    main()
