"""Phase timing of the first pano_attn unit of CTA 0 (VLN_PANO_STAMPS=1): clock64 deltas from kernel entry.
usage: python tools/pano_stamps.py [B ...]"""
import ctypes as C
import os
import sys

import torch

os.environ.setdefault("VLN_PANO_STAMPS", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clvln_b200  # noqa: E402,F401
from clvln_b200 import ops, _lib  # noqa: E402
from clvln_b200.environ import world as W  # noqa: E402

dev = torch.device("cuda:0")
n_vp = 10567
g = torch.Generator(device=dev).manual_seed(1)
z = torch.zeros(1, device=dev, dtype=torch.int32)
tables = dict(table=W.make_table(n_vp, 1, dev), cand_vp=z, cand_view=z, cand_ang4=torch.zeros(1, device=dev), n_cand=z,
              next_hop=z, dist=torch.zeros(1, device=dev), sq_off=torch.zeros(1, device=dev, dtype=torch.int64), vp_local=z,
              loc4=torch.from_numpy(W.static_loc4()).to(dev), pose4=torch.from_numpy(W.pose4()).to(dev))
store = ops.FeatureStore(tables, dev)
L = _lib.lib()
L.vln_debug_pano_stamps.argtypes = [C.c_void_p]
names = ["init", "vectors ready", "phase1 done (this thread)", "phase1 sync", "softmax sync", "phase2 done", "unit end"]
for B in [int(a) for a in sys.argv[1:]] or [16, 64, 2048]:
    for bits in (False, True):
        rng = ops.Rng(1, dev)
        view = torch.randint(0, 36, (B,), device=dev, dtype=torch.int32, generator=g)
        q = torch.randn(B, 2176, device=dev) * 0.05
        attn = torch.empty(B, 36, device=dev)
        out = torch.empty(B, 2176, device=dev)
        mb = None
        if bits:
            mb = torch.empty((B * 36, 256), dtype=torch.uint8, device=dev)
            ops._call("vln_feature_mask_bits", ops._ptr(mb), B * 36, 1, 0.3, rng.ptr, 1, 0, ops._stream())
        for _ in range(3):
            vp = torch.randint(0, n_vp, (B,), device=dev, dtype=torch.int32, generator=g)
            ops._call("vln_pano_attn_ld", store.handle, ops._ptr(vp), ops._ptr(view), ops._ptr(store.loc4), ops._ptr(q), 2176,
                      ops._ptr(attn), None, 2176, ops._ptr(out), 2176, B, 0, 0.3 if bits else 0.0, rng.ptr, 1, ops._ptr(mb), 1,
                      ops._stream())
            torch.cuda.synchronize()
        buf = (C.c_ulonglong * 16)()
        L.vln_debug_pano_stamps(buf)
        d = [buf[i] - buf[0] for i in range(1, 8)]
        print(f"B={B} mask_bits={bits}: " + ", ".join(f"{n}={v}" for n, v in zip(names, d)) + f" | unit starts at {buf[8] - buf[0]}, index in hand at {buf[9] - buf[0]}, rows requested at {buf[10] - buf[0]}")
