"""Warm, in-graph kernel timing of one EnvDrop training iteration: replays the captured CUDA graph of
bench.py's step under torch.profiler (CUPTI activity records, no replay/serialisation as under ncu) and
prints per-kernel launches / total us / share / average, plus the idle time between kernels.
Run with VLN_PDL=0 for per-kernel durations (under programmatic dependent launch a kernel's record includes
the time it spends waiting for its predecessor, so the records overlap).
usage: [VLN_PDL=0] python tools/trace_step.py [out.md]"""
import os
import random
import sys
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    import clvln_b200  # noqa: F401
    from clvln_b200 import utils
    from clvln_b200.agent import build_agent
    from clvln_b200.engine.graphs import GraphedTrainStep
    from clvln_b200.environ import R2RBatch
    torch.backends.cuda.matmul.allow_tf32 = False
    world, items = bench.build_world(False, dev)
    cfg = utils.agent_cfg(os.environ.get("VLN_TRACE_AGENT", "ENVDROP"))
    random.seed(2020)
    env = R2RBatch(world, items, batch_size=int(os.environ.get("VLN_TRACE_BATCH", "64")), device=dev)
    torch.manual_seed(2020)
    agent = build_agent(cfg, utils.StubTokenizer(), dev)
    agent.env = env
    agent.train()
    agent.sync_every = 0
    step = GraphedTrainStep(cfg, agent)
    for _ in range(12):
        step()
    torch.cuda.synchronize()
    n_it = 4
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(n_it):
            step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    tot, cnt = defaultdict(float), defaultdict(int)
    spans = []
    durs = defaultdict(list)
    for e in evs:
        name = e.name.replace("(anonymous namespace)::", "").replace("void ", "").split("<")[0].split("(")[0]
        d = e.device_time if hasattr(e, "device_time") else e.cuda_time
        tot[name] += d
        cnt[name] += 1
        durs[name].append(d)
        spans.append((e.time_range.start, e.time_range.end))
    spans.sort()
    busy = sum(b - a for a, b in spans)
    wall = spans[-1][1] - spans[0][0]
    total = sum(tot.values())
    lines = [f"{n_it} iterations: wall {wall / n_it:.0f} us/iteration, kernels busy {busy / n_it:.0f} us/iteration, "
             f"idle between kernels {(wall - busy) / n_it:.0f} us/iteration, {len(evs) / n_it:.0f} device activities/iteration", "",
             "| kernel | launches/iter | total us/iter | share | avg us |", "|---|---:|---:|---:|---:|"]
    for k in sorted(tot, key=tot.get, reverse=True)[:40]:
        lines.append(f"| {k[:70]} | {cnt[k] / n_it:.0f} | {tot[k] / n_it:.1f} | {100 * tot[k] / total:.1f}% | {tot[k] / cnt[k]:.2f} |")
    # duration histogram (1 us buckets) of the step-chain kernels: separates the GEMM shapes
    for k in ("linear_bf16x3_kernel", "pano_attn_kernel", "ctx_attn_kernel"):
        if k in durs:
            h = defaultdict(int)
            for d in durs[k]:
                h[int(d)] += 1
            lines.append("")
            lines.append(f"{k} durations (us bucket: launches/iter): " +
                         ", ".join(f"{b}-{b + 1}: {h[b] / n_it:.0f}" for b in sorted(h)))
    # raw timeline of the first profiled iteration (start offset us, duration us, stream, kernel) for gap analysis
    tl = os.environ.get("VLN_TIMELINE")
    if tl:
        per = len(evs) // n_it
        rows = sorted((e.time_range.start, e.time_range.end - e.time_range.start,
                       e.name.replace("(anonymous namespace)::", "").replace("void ", "").split("<")[0].split("(")[0])
                      for e in evs)[:per + 8]
        t0 = rows[0][0]
        with open(tl, "w") as f:
            for a, d, nm in rows:
                f.write(f"{a - t0:.2f},{d:.2f},{nm[:60]}\n")
    text = "\n".join(lines)
    print(text)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text + "\n")


if __name__ == "__main__":
    main()
