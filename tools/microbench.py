"""Config-5 microbench: gather + panorama attention (and candidate logits) HBM sweep.

Random viewpoints over the full-size table (1.56 GB >> 126 MB L2) so every launch reads from
HBM.  Prints one JSON line per (B, split, mode): achieved algorithmic GB/s = B*147456 / time."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clvln_b200  # noqa: E402
from clvln_b200 import ops  # noqa: E402
from clvln_b200.environ import world as W  # noqa: E402


def timeit(fn, iters, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for s, e in evs:
        s.record()
        fn()
        e.record()
    torch.cuda.synchronize()
    ts = sorted(s.elapsed_time(e) for s, e in evs)
    return ts[len(ts) // 2] * 1e-3, ts[0] * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-vp", type=int, default=10567)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--batches", type=int, nargs="*", default=[16, 64, 128, 256, 512, 1024, 2048])
    ap.add_argument("--splits", type=int, nargs="*", default=[1, 2, 4, 8])
    ap.add_argument("--drop", type=float, default=0.0)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    peaks = {}
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    peak = peaks.get("hbm_gbs", 6650.0)
    # tables: only what the kernels index (random candidate tables are enough for bandwidth)
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
    for B in args.batches:
        vp = torch.randint(0, n_vp, (B,), device=dev, dtype=torch.int32, generator=g)
        view = torch.randint(0, 36, (B,), device=dev, dtype=torch.int32, generator=g)
        q = torch.randn(B, 2176, device=dev) * 0.05
        attn = torch.empty(B, 36, device=dev)
        for split in args.splits:
            for mode in (0, 1):
                def run():
                    # fresh random viewpoints each launch would need a sync; the table is 12x L2 and B*147KB
                    # of it is touched per launch, so re-launching on the same indices at B<=512 can hit L2:
                    # rotate through 16 index sets.
                    run.k = (run.k + 1) % 16
                    ops.pano_attn_raw(store, vps[run.k], view, q, attn, mode, args.drop, 1, 2, split)
                vps = [torch.randint(0, n_vp, (B,), device=dev, dtype=torch.int32, generator=g) for _ in range(16)]
                run.k = 0
                med, best = timeit(run, args.iters)
                gbs = B * 147456 / med / 1e9
                print(json.dumps(dict(kernel="pano_attn", mode="fwd" if mode == 0 else "bwd", B=B, split=split,
                                      drop=args.drop, us=round(med * 1e6, 2), best_us=round(best * 1e6, 2),
                                      algo_GBs=round(gbs, 1), frac_of_measured_peak=round(gbs / peak, 3))), flush=True)
        tgt = torch.randn(B, 2176, device=dev) * 0.05
        logits = torch.empty(B, 16, device=dev)
        ncs = tables["n_cand"][vp.long()].sum().item()

        def run_c():
            ops._lib.check(ops._lib.lib().vln_cand_logits_fwd(
                store.handle, ops._ptr(vp), ops._ptr(view), ops._ptr(store.cand_view), ops._ptr(store.cand_ang4),
                ops._ptr(store.n_cand), ops._ptr(tgt), None, ops._ptr(logits), B, args.drop, 1, 2, ops._stream()))
        med, best = timeit(run_c, args.iters)
        print(json.dumps(dict(kernel="cand_logits_fwd", B=B, us=round(med * 1e6, 2),
                              algo_GBs=round(ncs * 4096 / med / 1e9, 1), note="same indices every launch (L2-warm)")),
              flush=True)


if __name__ == "__main__":
    main()
