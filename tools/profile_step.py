"""One EnvDrop training iteration between cudaProfilerStart/Stop, for
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ...
(the launch list committed under profiles/), or --set full -k regex:<kernel> captures."""
import argparse
import os
import random
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--warm", type=int, default=2)
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--agent", default="ENVDROP")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    import clvln_b200  # noqa: F401
    from clvln_b200 import utils
    from clvln_b200.agent import build_agent
    from clvln_b200.engine import TrainStep
    from clvln_b200.environ import R2RBatch
    torch.backends.cuda.matmul.allow_tf32 = False
    world, items = bench.build_world(args.small, dev)
    cfg = utils.agent_cfg(args.agent)
    random.seed(2020)
    env = R2RBatch(world, items, batch_size=args.batch, device=dev)
    torch.manual_seed(2020)
    agent = build_agent(cfg, utils.StubTokenizer(), dev)
    agent.env = env
    agent.train()
    agent.sync_every = 0
    step = TrainStep(cfg, agent)
    for _ in range(args.warm):
        step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
