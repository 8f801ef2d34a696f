"""Where a decoder step's time goes, on ONE clock across kernels (debug build of the library).

    VLN_LIB_VARIANT=stamps VLN_GEMM_STAMPS=1 python tools/chain_stamps.py [out.csv]

Replays the captured CUDA graph of bench.py's iteration with the step-chain kernels recording %globaltimer for block 0 /
thread 0: first instruction (CTA resident), after griddepcontrol.wait (predecessor complete + flushed), last
instruction.  Prints, for a window of forward and backward decoder steps, each kernel's
    resident-before-release (prologue overlapped with the predecessor),  release -> own end (work on the chain),
    own end -> successor's release (drain of the other CTAs + completion / flush latency).
"""
import ctypes as C
import os
import random
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

KID = {1: "pano_fwd", 11: "pano_bwd", 3: "ctx_step_fwd", 13: "ctx_step_bwd", 6: "tail", 100: "gemm"}


def main():
    assert os.environ.get("VLN_LIB_VARIANT") == "stamps", "run with VLN_LIB_VARIANT=stamps (debug build)"
    os.environ.setdefault("VLN_GEMM_STAMPS", "1")
    dev = torch.device("cuda:0")
    import clvln_b200  # noqa: F401
    from clvln_b200 import _lib, utils
    _lib.build()
    from clvln_b200.agent import build_agent
    from clvln_b200.engine.graphs import GraphedTrainStep
    from clvln_b200.environ import R2RBatch
    torch.backends.cuda.matmul.allow_tf32 = False
    world, items = bench.build_world(False, dev)
    cfg = utils.agent_cfg("ENVDROP")
    random.seed(2020)
    env = R2RBatch(world, items, batch_size=64, device=dev)
    torch.manual_seed(2020)
    agent = build_agent(cfg, utils.StubTokenizer(), dev)
    agent.env = env
    agent.train()
    agent.sync_every = 0
    step = GraphedTrainStep(cfg, agent)
    for _ in range(10):
        step()
    torch.cuda.synchronize()
    st = agent.rng.state
    st[3] = 0
    st[2] = 1                                   # enable
    L = _lib.lib()
    L.vln_debug_gemm_stamps.argtypes = [C.c_void_p, C.c_void_p]
    buf = (C.c_ulonglong * (1024 * 12))()
    n0 = C.c_uint()
    L.vln_debug_gemm_stamps(buf, C.byref(n0))
    step()
    torch.cuda.synchronize()
    st[2] = 0
    n1 = C.c_uint()
    L.vln_debug_gemm_stamps(buf, C.byref(n1))
    recs = []
    cnt = int(st[3].item())
    raw = st[4:4 + 4 * min(cnt, 8192)].cpu().view(-1, 4).tolist()
    for kid, t0, tw, te in raw:
        recs.append((t0, tw, te, KID.get(kid, str(kid))))
    for k in range(n0.value, n1.value):
        s = buf[(k % 1024) * 12:(k % 1024) * 12 + 12]
        ph = [s[j + 1] - s[j] if s[j + 1] >= s[j] and s[j] else 0 for j in range(8)]
        # cycles (1.965 GHz): setup | wait+first tiles | convert | mma | tmem->smem | sync | reductions(+epilogue) | teardown
        recs.append((s[10], s[9], s[11], "gemm[" + " ".join(str(int(c)) for c in ph) + "]"))
    recs = [r for r in recs if r[2] > 0]
    recs.sort(key=lambda r: r[2])
    base = recs[0][0]
    lines = ["start_us,release_us,end_us,kernel,resident_before_release,release_to_end,end_to_next_release"]
    for i, (t0, tw, te, name) in enumerate(recs):
        nxt = recs[i + 1][1] if i + 1 < len(recs) and recs[i + 1][1] > 0 else 0
        lines.append("%.2f,%.2f,%.2f,%s,%.2f,%.2f,%.2f" % ((t0 - base) / 1e3, (tw - base) / 1e3 if tw else -1, (te - base) / 1e3, name,
                                                        (tw - t0) / 1e3 if tw else -1, (te - tw) / 1e3 if tw else -1,
                                                        (nxt - te) / 1e3 if nxt else -1))
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/chain_stamps.csv"
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:1] + lines[200:260]))


if __name__ == "__main__":
    main()
