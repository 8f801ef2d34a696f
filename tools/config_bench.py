"""Throughput of the BASELINE.json configurations that are parity-test cases rather than the bench line, at full size
on the B200 (one JSON line each; run under torchrun for the data-parallel ones):

  config 2  EnvDrop, student forcing (sample), B=64                      python tools/config_bench.py --agent ENVDROP
  config 3  Self-Monitor + TRAIN.CLMODE=NAIVE, B=64/GPU                  ... --agent SELF-MONITOR --clmode NAIVE
  config 4  EnvDrop + TRAIN.CLMODE=SELF-PACE, B=128/GPU, full table      ... --agent ENVDROP --clmode SELF-PACE --batch 128
  (config 1, Follower teacher forcing, B=16)                             ... --agent FOLLOWER --batch 16

Each step is one full training iteration through the trainer's own step object (build_trainer(...).make_step): the
CUDA-graph replay of the fused EnvDrop rollout, the eager module path for Follower / Self-Monitor.  Device time with
CUDA events after warm-up, barrier + synchronize on both sides, max over ranks."""
import argparse
import json
import os
import random
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--agent", default="ENVDROP", choices=["ENVDROP", "FOLLOWER", "SELF-MONITOR"])
    ap.add_argument("--clmode", default="", choices=["", "NAIVE", "SELF-PACE"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--small", action="store_true")
    args = ap.parse_args()
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world_size > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        import datetime
        # a collective that does not complete aborts the job after 3 minutes instead of NCCL's default 10
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    import clvln_b200  # noqa: F401
    from clvln_b200 import utils
    from clvln_b200.agent import build_agent
    from clvln_b200.engine import build_trainer
    from clvln_b200.environ import CLR2RBatch, R2RBatch, split_rounds
    torch.backends.cuda.matmul.allow_tf32 = False
    world, items = bench.build_world(args.small, dev)
    cfg = utils.agent_cfg(args.agent)
    cfg.TRAIN.BATCH_SIZE = args.batch
    cfg.TRAIN.CLMODE = args.clmode
    random.seed(2020)
    if args.clmode == "SELF-PACE":          # configs/envdrop/envdrop_cl_config.yaml:31-39
        sp = cfg.TRAIN.SELF_PACE
        sp.FUNC, sp.LAMB, sp.MIU, sp.WCTRL, sp.CRATE, sp.INTERVAL, sp.BURN_IN, sp.STRATEGY = "linear", 2.0, 2.0, 0.5, 1.0, 10, 10, "epoch"
        env = CLR2RBatch(world, split_rounds(items), batch_size=args.batch, c_rate=sp.CRATE, device=dev, rank=rank,
                         world_size=world_size)
        train_env = env
    elif args.clmode == "NAIVE":            # round 1..k cumulative envs; the first round's env is what epoch 1 trains on
        rounds = split_rounds(items)
        train_env = {f"round_{k}": R2RBatch(world, [it for j in range(1, k + 1) for it in rounds[j]], batch_size=args.batch,
                                            device=dev, rank=rank, world_size=world_size) for k in range(1, 6)}
        env = train_env["round_3"]
    else:
        env = R2RBatch(world, items, batch_size=args.batch, device=dev, rank=rank, world_size=world_size)
        train_env = env
    torch.manual_seed(2020)
    agent = build_agent(cfg, utils.StubTokenizer(), dev)
    agent.env = env
    agent.train()
    trainer = build_trainer(cfg, train_env, dev)
    step = trainer.make_step(cfg, agent)
    if hasattr(step, "prefetch_next"):
        step.prefetch_next = True

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev)
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        n_ep = args.batch * world_size * args.steps
        print(json.dumps({"config": f"{args.agent} CLMODE={args.clmode or 'none'} B={args.batch}/GPU", "n_gpus": world_size,
                          "episodes_per_s": round(n_ep / float(t), 1), "ms_per_iteration": round(float(t) / args.steps * 1e3, 3),
                          "step": type(step).__name__, "loss": float(loss), "steps": args.steps,
                          "table_viewpoints": world.n_vp}), flush=True)
    if world_size > 1:                      # (iteration graphs hold captured NCCL kernels: no graceful communicator teardown)
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
