"""Data-parallel semantics on real GPUs: the gradient the optimiser applies under `torchrun` (one NCCL
all-reduce of the flat buffer, 1/world folded into the update kernel) equals the MEAN of the gradients a single
process computes on the ranks' shards of the same global minibatch (rows rank::world of the sorted 2B batch).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 tools/dp_check.py

Dropout off and teacher forcing, so every gradient is a deterministic function of (weights, shard).  Rank 0 prints one
JSON line: cosine / max-rel error between the all-reduced gradient / world and the single-process mean, whether the
parameters of all ranks are still identical after three optimiser steps, and the loss trajectory."""
import json
import os
import random
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def make(rank, world_size, dev, B):
    import clvln_b200  # noqa: F401
    from clvln_b200 import utils
    from clvln_b200.agent import build_agent
    from clvln_b200.engine import TrainStep
    from clvln_b200.environ import make_world, make_items, R2RBatch
    world = make_world(n_scans=4, seed=5, device=dev)
    items = make_items(world, 256, seed=5)
    cfg = utils.agent_cfg("ENVDROP")
    cfg.TRAIN.BATCH_SIZE = B
    cfg.AGENT.FEEDBACK = "teacher"
    cfg.AGENT.MAX_EPISODE_LEN = 12
    random.seed(2020)
    env = R2RBatch(world, items, batch_size=B, device=dev, rank=rank, world_size=world_size)
    torch.manual_seed(2020)
    agent = build_agent(cfg, utils.StubTokenizer(), dev)
    agent.env = env
    agent.train()
    agent.encoder.drop_ratio = agent.decoder.drop_ratio = agent.decoder.feat_drop_ratio = 0.0
    agent.critic.state2value[2].p = 0.0
    return cfg, env, agent, TrainStep(cfg, agent)


def main():
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    assert ws > 1, "run under torchrun with at least 2 ranks"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    B = 8
    # ---- data-parallel run: this rank's shard, gradient all-reduced by the optimiser's own call ----
    cfg, env, agent, step = make(rank, ws, dev, B)
    hook, agent._fused.on_grads_ready = agent._fused.on_grads_ready, None      # first without the early bucket: local gradient
    step.opt.zero_grad()
    loss, _ = step.losses()
    loss.backward()
    torch.cuda.synchronize()
    g_local = step.opt.grad.clone()
    # the same minibatch again through the optimiser's own path: decoder + critic bucket all-reduced from inside the
    # backward pass (under the encoder's BPTT), the encoder's bucket by finish_reduce
    agent._fused.on_grads_ready = hook
    early = os.environ.get("VLN_EARLY_ALLREDUCE", "1") != "0"
    assert (hook is not None) == early, "early all-reduce hook: installed exactly when enabled"
    env.reset_index = (lambda orig: (lambda **kw: orig(restart=True)))(env.reset_index)
    step.opt.zero_grad()
    loss, _ = step.losses()
    loss.backward()
    step.opt.finish_reduce()
    torch.cuda.synchronize()
    g_dp = step.opt.grad / ws
    g_sum = g_local.clone()
    dist.all_reduce(g_sum, op=dist.ReduceOp.SUM)
    bucket_err = float((g_dp - g_sum / ws).abs().max() / (g_sum / ws).abs().max())
    shard_ids = [it["instr_id"] for it in env.batch]
    out = {}
    if rank == 0:
        # ---- single process: the same global minibatch, each rank's shard in turn, mean of the gradients ----
        cfg1, env1, agent1, step1 = make(0, 1, dev, B * ws)
        agent1._fused.on_grads_ready = None                 # single-process reference: no collective may be issued here
        env1._next_minibatch()
        glob = list(env1.batch)
        assert [it["instr_id"] for it in glob[0::ws]] == shard_ids, "shard 0 is not rows 0::world of the sorted global batch"
        g_sum = torch.zeros_like(step1.opt.grad)
        for r in range(ws):
            env1._staged = []
            shard = glob[r::ws]
            orig = env1.reset_index
            env1.reset_index = lambda shard=shard, orig=orig, **kw: orig(batch=shard)
            step1.opt.zero_grad()
            l1, _ = step1.losses()
            l1.backward()
            torch.cuda.synchronize()
            env1.reset_index = orig
            g_sum += step1.opt.grad
        g_ref = g_sum / ws
        cos = float(torch.dot(g_dp, g_ref) / (g_dp.norm() * g_ref.norm()))
        rel = float((g_dp - g_ref).abs().max() / g_ref.abs().max())
        out.update(grad_cosine=cos, grad_max_rel=rel, bucketed_vs_single_allreduce_max_rel=bucket_err, local_vs_mean_cosine=float(torch.dot(g_local, g_ref) / (g_local.norm() * g_ref.norm())))
    # ---- three real optimiser steps: parameters must stay identical on every rank ----
    cfg, env, agent, step = make(rank, ws, dev, B)
    losses = [float(step()) for _ in range(3)]
    flat = step.opt.flat.clone()
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    same = torch.tensor([float(torch.equal(flat, ref))], device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    differing = {}
    if rank == 1:
        o = 0
        names = [(m_name, n_, p_) for m_name, m in zip(("encoder", "decoder", "critic"), agent._modules()) for n_, p_ in m.named_parameters()]
        for m_name, n_, p_ in names:
            off = (p_.data_ptr() - step.opt.flat.data_ptr()) // 4
            a, b = flat[off:off + p_.numel()], ref[off:off + p_.numel()]
            if not torch.equal(a, b):
                differing[f"{m_name}.{n_}"] = float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
        print("RANK1 differing parameters:", json.dumps(differing), flush=True)
    if rank == 0:
        out.update(world_size=ws, params_identical_after_3_steps=bool(same.item()), rank0_losses=[round(x, 4) for x in losses])
        print(json.dumps(out), flush=True)
        assert out["grad_cosine"] > 0.99999 and out["grad_max_rel"] < 1e-3 and out["params_identical_after_3_steps"]
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
