"""Experiment: does the GPU interleave two INDEPENDENT latency-bound iteration chains?
Two agents (own weights, own CUDA graph) with B/2 episodes each are replayed on two streams and timed against one
agent with B episodes.  If every dependent stage of the decoder chain costs ~5 us of mostly idle latency, two chains
should overlap and the pair should finish in well under 2x one chain's time.
usage: python tools/exp_two_chains.py [B] [n_chains]"""
import os
import random
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def make(world, items, B, dev, seed):
    from clvln_b200 import utils
    from clvln_b200.agent import build_agent
    from clvln_b200.engine.graphs import GraphedTrainStep
    from clvln_b200.environ import R2RBatch
    cfg = utils.agent_cfg("ENVDROP")
    cfg.TRAIN.BATCH_SIZE = B
    random.seed(seed)
    env = R2RBatch(world, items, batch_size=B, device=dev)
    torch.manual_seed(seed)
    agent = build_agent(cfg, utils.StubTokenizer(), dev)
    agent.env = env
    agent.train()
    agent.sync_every = 0
    return GraphedTrainStep(cfg, agent)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    n_chains = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    dev = torch.device("cuda:0")
    import clvln_b200  # noqa: F401
    torch.backends.cuda.matmul.allow_tf32 = False
    world, items = bench.build_world(False, dev)

    def timed(steps, streams, n=20):
        for _ in range(8):
            for st, s in zip(steps, streams):
                with torch.cuda.stream(s):
                    st()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            for st, s in zip(steps, streams):
                s.wait_stream(torch.cuda.current_stream()) if _ == 0 else None
                with torch.cuda.stream(s):
                    st()
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    one = make(world, items, B, dev, 2020)
    t1 = timed([one], [torch.cuda.Stream()])
    print(f"one chain, B={B}: {t1:.3f} ms/iteration = {B / t1 * 1e3:.0f} episodes/s", flush=True)
    del one
    half = B // n_chains
    steps = [make(world, items, half, dev, 2020 + k) for k in range(n_chains)]
    th = timed(steps[:1], [torch.cuda.Stream()])
    print(f"one chain, B={half}: {th:.3f} ms/iteration = {half / th * 1e3:.0f} episodes/s", flush=True)
    t2 = timed(steps, [torch.cuda.Stream() for _ in steps])
    print(f"{n_chains} chains x B={half} on {n_chains} streams: {t2:.3f} ms per round = {B / t2 * 1e3:.0f} episodes/s", flush=True)


if __name__ == "__main__":
    main()
