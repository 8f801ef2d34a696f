"""Phase timing of one encoder-LSTM timestep (VLN_LSTM_STAMPS=1): clock64 deltas inside step 10 of CTA (0,0)."""
import ctypes as C
import os
import sys

import torch

os.environ["VLN_LSTM_STAMPS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clvln_b200  # noqa: E402,F401
from clvln_b200 import ops, _lib  # noqa: E402

dev = torch.device("cuda:0")
B, L, H = 64, 80, 256
xproj = [torch.randn(B, L, 4 * H, device=dev) * 0.1 for _ in range(2)]
whh = [torch.randn(4 * H, H, device=dev) * 0.05 for _ in range(2)]
lengths = torch.full((B,), L, dtype=torch.int32, device=dev)
for _ in range(3):
    ops.lstm_layer(xproj, whh, lengths)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.lstm_layer(xproj, whh, lengths)
e1.record()
torch.cuda.synchronize()
print("fwd launch: %.1f us (%.2f us per timestep)" % (e0.elapsed_time(e1) * 100, e0.elapsed_time(e1) * 100 / L))
L_ = _lib.lib()
L_.vln_debug_lstm_stamps.argtypes = [C.c_void_p]
buf = (C.c_ulonglong * 12)()
L_.vln_debug_lstm_stamps(buf)
names = ["mma + gate stores", "syncthreads", "pointwise + global stores", "st.async issue", "mbarrier wait"]
if buf[0]:           # stamps of the mma.sync kernels (VLN_LSTM_VARIANT=mma)
    print(", ".join(f"{n}={buf[i + 1] - buf[i]}" for i, n in enumerate(names)), "cycles; step total", buf[5] - buf[0])
    print("kernel: prologue (weight fragments, barrier init, cluster sync) = %d cycles, %d steps = %d cycles (%.0f per step), epilogue = %d"
          % (buf[7] - buf[6], L, buf[8] - buf[7], (buf[8] - buf[7]) / L, buf[9] - buf[8]))
L_.vln_debug_lstm_tc_stamps.argtypes = [C.c_void_p]
tb = (C.c_ulonglong * 16)()
L_.vln_debug_lstm_tc_stamps(tb)
if tb[0]:
    print("tcgen05 fwd step 10 (cycles from loop top): h arrived=%d, MMAs issued=%d, accumulator ready=%d, TMEM->smem=%d, "
          "syncthreads=%d, cell update=%d, sends issued=%d" % tuple(int(tb[i]) - int(tb[0]) for i in range(1, 8)))
    print("tcgen05 fwd kernel: zero+alloc+weights->TMEM=%d, cluster sync=%d, %d steps=%d (%.0f per step)"
          % (tb[9] - tb[8], tb[10] - tb[9], L, tb[11] - tb[10], (tb[11] - tb[10]) / L))
L_.vln_debug_lstm_occupancy.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
f, b = C.c_int(0), C.c_int(0)
for h in (256, 128):
    if L_.vln_debug_lstm_occupancy(h, C.byref(f), C.byref(b)) == 0:
        print(f"H={h}: max resident clusters fwd={f.value} bwd={b.value} (mma.sync kernels, 8 rows per cluster: a B=64 "
              f"bidirectional launch needs 16; the tcgen05 kernels take 16 / 24 / 32 rows per cluster)")
# one wave or two?  time the launch at 8, 16, 24, 32 clusters (B = 32 .. 128, both directions)
for Bx in (32, 56, 64, 96, 128):
    xp = [torch.randn(Bx, L, 4 * H, device=dev) * 0.1 for _ in range(2)]
    ln = torch.full((Bx,), L, dtype=torch.int32, device=dev)
    for _ in range(2):
        ops.lstm_layer(xp, whh, ln)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        ops.lstm_layer(xp, whh, ln)
    e1.record()
    torch.cuda.synchronize()
    print("B=%d, both directions: fwd launch %.1f us" % (Bx, e0.elapsed_time(e1) * 100))
