"""Microbenchmark of the skinny tcgen05 linear (vln_linear_bf16x3) on the ten GEMM shapes of one EnvDrop
decoder step (five forward, five input-gradient), warm L2, launched back to back inside a CUDA graph
the way the rollout runs them.  Prints one JSON line per shape: us per launch, weight bytes streamed
(bf16 hi + lo), achieved GB/s and TFLOP/s (3 MMAs per product)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clvln_b200  # noqa: E402,F401
from clvln_b200 import ops  # noqa: E402
from clvln_b200.agent.fused import _gemm, _p  # noqa: E402

SHAPES = [("q = W_vin h", 2176, 512), ("gates = Wcat xh", 2048, 2752), ("tq = W_tin h", 512, 512),
          ("pre = W_out wh", 512, 1024), ("tgt = W_cand h", 2176, 512),
          ("d_hc = dtgt W_cand", 512, 2176), ("d_wh = dpre W_out", 1024, 512), ("d_h = dtq W_tin", 512, 512),
          ("d_xh = dgates Wcat", 2752, 2048), ("d_hq = dq W_vin", 512, 2176)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    M = args.batch
    # launch floor: a one-thread kernel, back to back in the same kind of graph
    st8 = torch.zeros(2, dtype=torch.int64, device=dev)
    g0 = torch.cuda.CUDAGraph()
    ops._call("vln_rng_advance", ops._ptr(st8), 0, ops._stream())
    torch.cuda.synchronize()
    with torch.cuda.graph(g0):
        for _ in range(16):
            ops._call("vln_rng_advance", ops._ptr(st8), 0, ops._stream())
    g0.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        g0.replay()
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"gemm": "(empty kernel)", "us": round(e0.elapsed_time(e1) * 1e3 / (args.reps * 16), 2),
                      "variant": os.environ.get("VLN_GEMM_VARIANT", "c8")}), flush=True)
    for name, N, K in SHAPES:
        if args.only and args.only not in name:
            continue
        w = torch.randn(N, K, device=dev) * 0.02
        sw = ops._SplitWeight(w).fresh(w)
        x = torch.randn(M, K, device=dev)
        y = torch.zeros(M, N, device=dev)
        n_in = 16
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                _gemm(sw.hi, sw.lo, N, K, _p(x), K, M, None, _p(y), N)
        torch.cuda.current_stream().wait_stream(s)
        ref = x.double() @ w.double().t()
        y.zero_()
        _gemm(sw.hi, sw.lo, N, K, _p(x), K, M, None, _p(y), N)
        err = float((y.double() - ref).abs().max() / ref.abs().max())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(n_in):
                _gemm(sw.hi, sw.lo, N, K, _p(x), K, M, None, _p(y), N)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (args.reps * n_in)
        wbytes = N * K * 4
        print(json.dumps({"gemm": name, "M": M, "N": N, "K": K, "us": round(us, 2), "max_rel_err": float("%.2e" % err),
                          "weight_GBs": round(wbytes / us / 1e3, 1),
                          "TFLOPs_bf16x3": round(3 * 2 * M * N * K / us / 1e6, 2)}), flush=True)


def stamps():
    import ctypes as C
    from clvln_b200 import _lib
    L = _lib.lib()
    buf = (C.c_ulonglong * (1024 * 12))()
    n = C.c_uint()
    L.vln_debug_gemm_stamps.argtypes = [C.c_void_p, C.c_void_p]
    L.vln_debug_gemm_stamps(buf, C.byref(n))
    k = min(n.value, 64)
    rows = [[buf[i * 12 + j] for j in range(12)] for i in range(k)]
    rows.sort(key=lambda r: r[10])
    print("launches recorded:", n.value)
    prev_end = None
    for r in rows[-12:]:
        d = [r[j] - r[0] for j in range(1, 9)]
        gap = (r[10] - prev_end) if prev_end else 0
        print("cycles since entry: prologue=%d w_landed=%d x_ready=%d acc_done=%d tile_in_smem=%d synced=%d stored=%d end=%d | wall %d ns, gap from previous end %d ns"
              % (*d, r[11] - r[10], gap))
        prev_end = r[11]


if __name__ == "__main__":
    main()
    if os.environ.get("VLN_GEMM_STAMPS"):
        stamps()
