#!/usr/bin/env python
"""Headline benchmark: EnvDrop training iterations per second (episodes/s), B=64 per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...    # reference algorithm on host cores

One "step" = one full EnvDrop training iteration of the reference (trainer.py:411-429): a
teacher-forced rollout (imitation loss) + a sampled 35-step rollout on the same minibatch (A2C) +
backward + clip(encoder, 40) + clip(decoder, 40) + RMSprop, on B=64 episodes per GPU with
instruction length 80, 36 views x 2176 features, hidden 512 (BASELINE.json configs[1]).
Synthetic data of the real shape: the full 10 567-viewpoint bf16 feature table (1.56 GB, HBM
resident), 90 random connectivity graphs, 14 039 episodes; random-init weights.

The line printed by rank 0 carries, besides the base contract:
  value     episodes/s with the minibatch index tensors staged in HBM before the timed region;
  e2e       the same through the public TrainStep API: per-step pinned host->device copies of the
            minibatch (tokens, lengths, start pose, goal) and a device->host read of the loss;
  roofline  the fused gather + 36-view attention kernel, timed alone with CUDA events (a CUDA graph
            of launches over rotating random viewpoints), algorithmic bytes B*36*2048*2 per launch;
  cpu_baseline  the oracle's CPU restatement of the reference iteration on this host's cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train episodes/sec (EnvDrop IL+A2C iteration, B=64/GPU, L=80)"
UNIT = "episodes/s"
B_PER_GPU = 64
WORKLOAD = ("EnvDrop IL+A2C training iteration (teacher rollout + 35-step sampled rollout + backward + clip + RMSprop), "
            "B=%d/GPU, L=80, 36x2176 features, H=512")
ALGO_BYTES_PER_EPISODE_STEP = 36 * 2048 * 2


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU)
    ap.add_argument("--small", action="store_true", help="tiny world (debug)")
    ap.add_argument("--lengths", default="80", choices=["80", "real"],
                    help="instruction lengths: 80 for every row (headline, worst case) or drawn from the real R2R "
                         "training distribution (mean 31.3; SURVEY 8d's second run)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--per-op", action="store_true",
                    help="with --impl reference: add per-op CPU times (gather, panorama attention, LSTMCell, context "
                         "attention, cross-entropy; SURVEY 8d) and the config-1 Follower iteration to the line")
    ap.add_argument("--graph", type=int, default=int(os.environ.get("VLN_BENCH_GRAPH", "1")),
                    help="replay the iteration as CUDA graphs (1) or launch eagerly (0)")
    return ap.parse_args()


def build_world(small, device, with_table=True, lengths="80"):
    import clvln_b200  # noqa: F401
    from clvln_b200.environ import make_world, make_items, full_world_sizes
    from clvln_b200.environ.world import r2r_length_counts
    kw = dict(fixed_len=80) if lengths == "80" else dict(length_counts=r2r_length_counts())
    if small:
        world = make_world(n_scans=6, seed=2020, device=device, with_table=with_table)
        items = make_items(world, 2000, seed=2020, **kw)
    else:
        world = make_world(sizes=full_world_sizes(), seed=2020, device=device, with_table=with_table)
        items = make_items(world, 14039, seed=2020, **kw)
    return world, items


# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (profiling recipe's line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.p, self.rows = gpu_index, None, []

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower() == "active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def roofline_pano(store, ops, torch, B, split, peaks):
    """Time the fused gather+attention kernel alone, configured exactly as the rollout launches it
    (feature dropout 0.3 through pre-generated keep-bits, output into the strided LSTM operand row):
    a CUDA graph of 32 launches over rotating random viewpoint sets so that every launch reads HBM,
    not L2; CUDA events on the launch stream."""
    import ctypes as C
    dev = store.device
    g = torch.Generator(device=dev).manual_seed(7)
    n_sets = 32
    vps = [torch.randint(0, store.n_vp, (B,), device=dev, dtype=torch.int32, generator=g) for _ in range(n_sets)]
    view = torch.randint(0, 36, (B,), device=dev, dtype=torch.int32, generator=g)
    q = torch.randn(B, 2176, device=dev) * 0.05
    attn = torch.empty(B, 36, device=dev)
    out = torch.empty(B, 2752, device=dev)
    rng = ops.Rng(1, dev)
    # as the rollout launches it: pre-generated packed keep-bits are streamed next to the rows
    use_bits = True
    bits = None
    if use_bits:
        bits = torch.empty((n_sets, B * 36, 256), dtype=torch.uint8, device=dev)
        ops._call("vln_feature_mask_bits", ops._ptr(bits), B * 36, n_sets, 0.3, rng.ptr, 1, 7, ops._stream())

    def launch(k, mode=0):
        ops._call("vln_pano_attn_ld", store.handle, ops._ptr(vps[k]), ops._ptr(view), ops._ptr(store.loc4), ops._ptr(q),
                  2176, ops._ptr(attn), None, 2176, C.c_void_p(out.data_ptr() + 256), 2752, B, mode, 0.3, rng.ptr, 1 + 7 * k,
                  ops._ptr(bits[k]) if use_bits else None, split, ops._stream())
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for k in range(3):
            launch(k)
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for k in range(n_sets):
            launch(k)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3 / (reps * n_sets)
    # latency floor of a launch of this shape: the same kernel, same grid / cluster / shared memory / dependency chain,
    # stopping after its dependent index load and ONE row's HBM round trip per CTA (mode bit 2)
    floor = None
    if B <= torch.cuda.get_device_properties(dev).multi_processor_count // 2:
        gf = torch.cuda.CUDAGraph()
        launch(0, 4)
        torch.cuda.synchronize()
        with torch.cuda.graph(gf):
            for k in range(n_sets):
                launch(k, 4)
        for _ in range(3):
            gf.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            gf.replay()
        e1.record()
        torch.cuda.synchronize()
        floor = e0.elapsed_time(e1) * 1e-3 / (reps * n_sets)
    achieved = B * ALGO_BYTES_PER_EPISODE_STEP / t / 1e9
    peak = float(peaks.get("hbm_gbs", 6650.0))
    return {"bound": "hbm", "kernel": "pano_attn fwd (fused gather + feature dropout + 36-view soft-dot attention)",
            "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
            "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback",
            "peak_spec": 8000.0, "frac_of_spec": round(achieved / 8000.0, 4),      # both denominators
            "us_per_launch": round(t * 1e6, 2), "episodes_per_launch": B, "split": split, "traffic": None,
            **({"latency_floor_us": round(floor * 1e6, 2),
                "transfer_us_at_peak": round(B * ALGO_BYTES_PER_EPISODE_STEP / (float(peaks.get("hbm_gbs", 6650.0)) * 1e9) * 1e6, 2),
                "frac_of_floor_plus_transfer": round((floor + B * ALGO_BYTES_PER_EPISODE_STEP / (float(peaks.get("hbm_gbs", 6650.0)) * 1e9)) / t, 4),
                "floor_note": "latency_floor_us = this kernel's launch (same grid, 2-CTA clusters, 216 KB shared memory, same dependency chain) + the dependent index load + one row's HBM round trip per CTA, no payload; a launch cannot beat floor + bytes/peak"}
               if floor is not None else {}),
            "keep_bits": "pre-generated, streamed (9 216 B / episode)" if use_bits else "drawn in shared memory ahead of the dependency wait",
            "note": "B=64 episodes per launch is the north-star shape: 9.4 MB per launch = 1.4 us at peak, so the launch is latency-bound; tools/microbench.py sweeps B up to 2048"}


def roofline_others(ops, torch, B, peaks):
    """The other kernel families of the iteration, each timed alone in bench.py (CUDA graph of 16 back-to-back
    launches, CUDA events): the gates GEMM — the largest of the skinny tcgen05 linears that take a third of the
    iteration — against the tensor peak and as streamed weight bytes (L2-resident: 22.5 MB of bf16 hi+lo per launch),
    and the encoder recurrence (tcgen05, W_hh in tensor memory) as time per step of its 80-step latency chain."""
    from clvln_b200.agent.fused import _gemm, _p
    dev = torch.device("cuda", torch.cuda.current_device())
    out = []

    def time_graph(fn, n_in=16, reps=10):
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                fn()
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(n_in):
                fn()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / (reps * n_in)

    N, K = 2048, 2752
    w = torch.randn(N, K, device=dev) * 0.02
    sw = ops._SplitWeight(w).fresh(w)
    x = torch.randn(B, K, device=dev)
    y = torch.zeros(B, N, device=dev)
    t = time_graph(lambda: _gemm(sw.hi, sw.lo, N, K, _p(x), K, B, None, _p(y), N))
    tf = 3 * 2 * B * N * K / t / 1e12
    peak_tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1350.0)))
    out.append({"kernel": "linear_bf16x3 (LSTMCell gates, %dx%dx%d, bf16x3 = 3 MMAs per product)" % (N, K, B), "bound": "l2",
                "us_per_launch": round(t * 1e6, 2), "weight_GBs_from_l2": round(N * K * 4 / t / 1e9, 1),
                "achieved_tflops": round(tf, 1), "peak_tflops": peak_tf, "tensor_frac": round(tf / peak_tf, 4),
                "useful_tflops": round(tf / 3, 1), "tensor_frac_useful": round(tf / 3 / peak_tf, 4),
                "note": "skinny (M = batch): bound by streaming the L2-resident weights (+ the activation tile every N-tile CTA re-reads), not by the tensor pipe; tensor_frac counts the three bf16 MMAs of every bf16x3 product, tensor_frac_useful counts 2*M*N*K once"})
    L, H = 80, 256
    xproj = [torch.randn(B, L, 4 * H, device=dev) * 0.1 for _ in range(2)]
    whh = [torch.randn(4 * H, H, device=dev) * 0.05 for _ in range(2)]
    lengths = torch.full((B,), L, dtype=torch.int32, device=dev)
    t = time_graph(lambda: ops.lstm_layer(xproj, whh, lengths), n_in=4, reps=5)
    tf = 2 * 3 * 2 * B * L * 4 * H * H / t / 1e12
    out.append({"kernel": "lstm_tc_fwd (encoder BiLSTM recurrence, B=%d, L=%d, H=%d per direction)" % (B, L, H), "bound": "latency",
                "us_per_launch": round(t * 1e6, 1), "us_per_timestep": round(t * 1e6 / L, 3),
                "achieved_tflops": round(tf, 2), "peak_tflops": peak_tf, "tensor_frac": round(tf / peak_tf, 5),
                "note": "80 serial steps; per step and CTA 48 tcgen05.mma of 128x16x16 (8 cycles each) inside a ~2 900-cycle exchange / gate-math chain"})
    # text-attention stage of a decoder step (csrc/ctx_step.cu): LSTM pointwise + masked softmax over the instruction +
    # weighted context, one CTA per episode; algorithmic bytes = the episode's CW and ctx tiles, 2 x L x H x 4 B (L2-resident)
    H = 512
    gates, c0 = torch.randn(B, 4 * H, device=dev), torch.randn(B, H, device=dev)
    h1, c1, acts = torch.empty(B, H, device=dev), torch.empty(B, H, device=dev), torch.empty(B, 4 * H, device=dev)
    wh, attn = torch.zeros(B, 2 * H, device=dev), torch.empty(B, L, device=dev)
    ctx, cw = torch.randn(B, L, H, device=dev), torch.randn(B, L, H, device=dev) * 0.05
    rng = ops.Rng(3, dev)
    t = time_graph(lambda: ops._call("vln_envdrop_ctx_step_fwd", ops._ptr(gates), ops._ptr(c0), ops._ptr(h1), ops._ptr(c1),
                                     ops._ptr(acts), ops._ptr(wh), 2 * H, ops._ptr(ctx), ops._ptr(cw), ops._ptr(lengths),
                                     ops._ptr(attn), B, L, H, 0.5, rng.ptr, 1, 1, ops._stream()))
    by = B * 2 * L * H * 4
    out.append({"kernel": "ctx_step_fwd (LSTM pointwise + text attention, B=%d, L=%d, H=%d)" % (B, L, H), "bound": "l2",
                "us_per_launch": round(t * 1e6, 2), "algorithmic_bytes": by, "achieved_GBs": round(by / t / 1e9, 1),
                "note": "per-episode CTA: one bulk copy of the dotted tile into shared memory + the summed tile in registers; latency-bound at B=64 (21 MB per launch)"})
    # weight gradient of the LSTMCell gates over the stacked rows of the paired rollout (csrc/wgrad.cu: tcgen05 kind::tf32,
    # both operands MN-major straight from TMA): the largest of the backward pass's dY^T X products
    R, M, Nw = 35 * 2 * B, 2048, 2752
    dy, xs = torch.randn(R, M, device=dev), torch.randn(R, Nw, device=dev)
    t = time_graph(lambda: ops.wgrad_tc(dy, xs), n_in=4, reps=5)
    tf = 2 * R * M * Nw / t / 1e12
    out.append({"kernel": "wgrad_tf32 (dW = dY^T X of the LSTMCell gates, R=%d rows, %dx%d, tcgen05 kind::tf32)" % (R, M, Nw),
                "bound": "l2", "us_per_launch": round(t * 1e6, 1), "achieved_tflops": round(tf, 1),
                "peak_tflops_tf32": round(peak_tf / 2, 1), "tensor_frac_tf32": round(tf / (peak_tf / 2), 4),
                "operand_GBs_from_l2": round((M // 128) * ((Nw + 127) // 128) * R * 1024 / t / 1e9, 1),
                "note": "fp32 operands as tf32: a 128x128 block needs 1 KB of operands per reduction row for 16 384 MACs, so "
                        "the launch is bound by L2 -> shared-memory traffic (operand_GBs_from_l2), not by the tensor pipe; "
                        "tf32 peak taken as half the measured dense bf16 rate"})
    return out


def ncu_traffic(csv_name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the committed `ncu --set full` capture
    (profiles/, trimmed by tools/ncu_trim.py); None when the capture is not there."""
    import csv
    path = os.path.join(ROOT, "profiles", csv_name)
    if not os.path.exists(path):
        return None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for row in csv.reader(open(path)):
        if row and row[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            vals = [float(v) for v in row[2:]]
            tot += scale.get(row[1], 1.0) * sum(vals) / len(vals)
    return int(tot) if tot else None


def cpu_iteration_factory(world_small, items, B, threads):
    """The oracle's CPU restatement of one reference EnvDrop training iteration (test infrastructure
    used here only as the timed CPU baseline)."""
    import random
    import torch
    from clvln_b200 import utils
    from clvln_b200.model import EncoderLSTM, EnvDropDecoder, Critic
    from oracle import port_env as PE, port_modules as P, port_rollout as PR
    torch.set_num_threads(threads)
    cfg = utils.agent_cfg("ENVDROP")
    mc = cfg.MODEL.ENVDROP
    torch.manual_seed(2020)
    mods = [EncoderLSTM(992, mc.WORD_EMB_SIZE, mc.HIDDEN_SIZE, 0, mc.DROP_RATE, True, 1),
            EnvDropDecoder(mc.HIDDEN_SIZE, mc.DROP_RATE, mc.FEAT_DROP_RATE, mc.ACT_EMB_SIZE),
            Critic(mc.HIDDEN_SIZE, mc.DROP_RATE)]
    sds = [{k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()} for m in mods]
    params = [v for sd in sds for v in sd.values()]
    opt = torch.optim.RMSprop(params, lr=cfg.TRAIN.LR)
    random.seed(2020)
    penv = PE.R2RBatchPort(PE.WorldView(world_small), items, batch_size=B)
    ag = PR.Agent("ENVDROP", sds[0], sds[1], sds[2], hidden=512, bidirectional=True, enc_layers=1, episode_len=35)
    drop = P.Drop("torch")
    n_enc, n_dec = len(sds[0]), len(sds[1])

    def iteration():
        _, l1 = PR.rollout_envdrop(ag, penv, train_ml=True, train_rl=False, feedback="teacher", drop=drop)
        _, l2 = PR.rollout_envdrop(ag, penv, train_ml=False, train_rl=True, restart=True, feedback="sample", drop=drop)
        loss = l1["ml_loss"] + l2["rl_loss"]
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params[:n_enc], 40.0)
        torch.nn.utils.clip_grad_norm_(params[n_enc:n_enc + n_dec], 40.0)
        opt.step()
        return float(loss)
    return iteration


def reference_iteration_factory(world, items, B, threads):
    """One EnvDrop training iteration of the UNMODIFIED reference (trainer.py:411-429: EnvDropAgent.rollout teacher +
    sampled, backward, clip_grad_norm x2, RMSprop) on the host cores: the reference's own agent / decoder / R2RBatch code
    from oracle/_ref (oracle/build_ref.py), with only its external inputs substituted (oracle/ref_harness.py: the
    Matterport simulator -> a table-driven graph walker, dataset / connectivity loaders -> the synthetic world).
    Returns None when oracle/_ref is not there."""
    import contextlib
    import random
    import torch
    from oracle import ref_loader, ref_harness as H
    if not ref_loader.reference_available():
        return None
    torch.set_num_threads(threads)
    with contextlib.redirect_stdout(sys.stderr):           # the reference prints progress lines; stdout carries ONE json line
        src = H.install(world, {"train": items})
        import src.environ as environ
        import src.agent as agent_mod
        tok = H.StubTokenizer(items)
        fs = H.feature_store(world)
        random.seed(2020)
        torch.manual_seed(2020)
        env = environ.R2RBatch(fs, batch_size=B, splits=["train"], tokenizer=tok)
        cfg = H.model_cfg("ENVDROP")
        ag = agent_mod.EnvDropAgent(cfg, 80, "/tmp", torch.device("cpu"), env, tok, episode_len=35)
    ag.env = env
    ag.train()
    params = list(ag.encoder.parameters()) + list(ag.decoder.parameters()) + list(ag.critic.parameters())
    opt = torch.optim.RMSprop(params, lr=1e-4)

    def iteration():
        with contextlib.redirect_stdout(sys.stderr):
            ag.rollout(train_ml=True, train_rl=False, feedback="teacher")
            ml = ag.loss["ml_loss"]
            ag.rollout(train_ml=False, train_rl=True, restart=True, feedback="sample")
            loss = ml + ag.loss["rl_loss"]
            opt.zero_grad()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(ag.encoder.parameters(), 40.0)
            torch.nn.utils.clip_grad_norm_(ag.decoder.parameters(), 40.0)
            opt.step()
        return float(loss.detach())
    return iteration


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def cpu_per_op(B, threads, world, items):
    """Per-op host times of the reference algorithm (oracle restatement, fp32, all host threads), forward + backward,
    at the EnvDrop step's shapes; plus one Follower teacher-forcing iteration at B=16 (BASELINE config 1)."""
    import numpy as np
    import torch
    import torch.nn.functional as F
    from clvln_b200.environ.world import static_loc4
    from oracle import port_modules as P
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(0)

    def clock(fn, reps=5):
        fn()
        t0 = time.time()
        for _ in range(reps):
            fn()
        return (time.time() - t0) / reps * 1e3

    out = {}
    vp = torch.randint(0, world.n_vp, (B,), generator=g)
    view = torch.randint(0, 36, (B,), generator=g)
    loc = torch.from_numpy(static_loc4())
    table = world.table.float()

    def gather():            # R2RBatch.observe concat + _feature_variable (common_env.py:309, base.py:141-147)
        feats = [np.concatenate((table[int(v)].numpy(), np.repeat(loc[int(w)].numpy(), 32, axis=1)), 1) for v, w in zip(vp, view)]
        return torch.from_numpy(np.stack(feats))
    out["gather_pano_ms"] = clock(gather)
    img = gather()
    h = torch.randn(B, 512, generator=g, requires_grad=True)
    w_in = (torch.randn(2176, 512, generator=g) * 0.02).requires_grad_(True)

    def pano():
        o, _ = P.soft_dot_attention(h, img, w_in)
        o.sum().backward()
    out["pano_attention_fwd_bwd_ms"] = clock(pano)
    x = torch.randn(B, 2240, generator=g, requires_grad=True)
    c = torch.randn(B, 512, generator=g)
    w_ih = (torch.randn(2048, 2240, generator=g) * 0.02).requires_grad_(True)
    w_hh = (torch.randn(2048, 512, generator=g) * 0.02).requires_grad_(True)
    bz = torch.zeros(2048)

    def cell():
        h1, c1 = P.lstm_cell(x, h, c, w_ih, w_hh, bz, bz)
        (h1.sum() + c1.sum()).backward()
    out["lstm_cell_fwd_bwd_ms"] = clock(cell)
    ctx = torch.randn(B, 80, 512, generator=g, requires_grad=True)
    w_t = (torch.randn(512, 512, generator=g) * 0.02).requires_grad_(True)
    w_o = (torch.randn(512, 1024, generator=g) * 0.02).requires_grad_(True)

    def ctxa():
        o, _ = P.soft_dot_attention(h, ctx, w_t, w_o)
        o.sum().backward()
    out["ctx_attention_fwd_bwd_ms"] = clock(ctxa)
    logit = torch.randn(B, 16, generator=g, requires_grad=True)
    tgt = torch.randint(0, 16, (B,), generator=g)

    def ce():
        F.cross_entropy(logit, tgt, ignore_index=-1, reduction="sum").backward()
    out["cross_entropy_fwd_bwd_ms"] = clock(ce)
    # config 1: Follower, teacher forcing, B=16, MAX_EPISODE_LEN=10
    import random
    from clvln_b200 import utils
    from clvln_b200.model import EncoderLSTM, AttnDecoderLSTM
    from oracle import port_env as PE, port_rollout as PR
    cfg = utils.agent_cfg("FOLLOWER")
    mc = cfg.MODEL.FOLLOWER
    torch.manual_seed(2020)
    mods = [EncoderLSTM(992, mc.WORD_EMB_SIZE, mc.HIDDEN_SIZE, 0, mc.DROP_RATE, mc.ENC_BIDIRECTION, mc.ENC_LAYERS),
            AttnDecoderLSTM(mc.HIDDEN_SIZE, mc.DROP_RATE)]
    sds = [{k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()} for m in mods]
    params = [v for sd in sds for v in sd.values()]
    opt = torch.optim.Adam(params, lr=cfg.TRAIN.LR)
    random.seed(2020)
    penv = PE.R2RBatchPort(PE.WorldView(world), items, batch_size=16)
    ag = PR.Agent("FOLLOWER", sds[0], sds[1], None, hidden=mc.HIDDEN_SIZE, bidirectional=mc.ENC_BIDIRECTION,
                  enc_layers=mc.ENC_LAYERS, episode_len=10)
    drop = P.Drop("torch")

    def follower():
        _, ml = PR.rollout_follower(ag, penv, feedback="teacher", drop=drop)
        opt.zero_grad()
        ml.backward()
        opt.step()
    ms = clock(follower, reps=3)
    out["follower_B16_teacher_iteration_ms"] = ms
    out["follower_B16_episodes_per_s"] = 16 / (ms * 1e-3)
    return {k: round(v, 3) for k, v in out.items()}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the iteration on all host threads, same metric and
    workload — the UNMODIFIED reference agent from oracle/_ref when it is there (kind "reference": full 90-scan world,
    the same 14 039 synthetic episodes as the CUDA arm), else the oracle port (kind "port", 8-scan world)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    threads = host_threads()
    from oracle import ref_loader
    B = args.batch
    real = ref_loader.reference_available()
    if real:
        world, items = build_world(args.small, None, with_table=True, lengths="80")
        make = lambda b: reference_iteration_factory(world, items, b, threads)
        world_desc = "%d-viewpoint synthetic world (%d scans)" % (world.n_vp, len(world.scans))
    else:
        from clvln_b200.environ import make_world, make_items
        world = make_world(n_scans=8, seed=2020)
        items = make_items(world, max(4 * B, 256), seed=2020, fixed_len=80)
        make = lambda b: cpu_iteration_factory(world, items, b, threads)
        world_desc = "8-scan synthetic world"
    it = make(B)
    t0 = time.time()
    it()                                                  # probe (also the first warm-up)
    probe = time.time() - t0
    budget = 200.0
    n_total = args.steps + args.warmup
    if probe * n_total > budget and B > 16:               # bound the run: a step becomes a 16-episode sample of the batch
        B = 16
        it = make(B)
        it()
    for _ in range(max(0, args.warmup - 1)):
        it()
    t0 = time.time()
    for _ in range(args.steps):
        it()
    dt = time.time() - t0
    v = B * args.steps / dt
    what = "the unmodified reference (src/agent/envdrop.py rollout x2 + backward + clip_grad_norm x2 + RMSprop, oracle/_ref)" \
        if real else "oracle CPU restatement of the reference"
    sample = (f"{args.steps} full EnvDrop training iterations (teacher + 35-step sampled rollout + backward + clip + RMSprop) at "
              f"B={B}, L=80, {world_desc}, fp32, {threads} torch threads: {what}")
    extra = {"per_op_ms": cpu_per_op(args.batch, threads, world, items)} if args.per_op else {}
    print(json.dumps({**extra,
        "impl": "reference", "metric": METRIC, "value": round(v, 3), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD % B, "device": "host CPU: " + cpu_model(), "torch_threads": threads,
                   "world": world_desc},
        "cpu_baseline": {"value": round(v, 3), "unit": UNIT, "cores": threads, "cpu": cpu_model(),
                         "kind": "reference" if real else "port", "sample": sample},
        "e2e": {"value": round(v, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def run_b200(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world_size > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":      # keeps the version banner off stdout (one JSON line)
            os.environ["NCCL_DEBUG"] = "WARN"
        import datetime
        # a collective that does not complete aborts the job after 3 minutes instead of NCCL's default 10
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    import clvln_b200  # noqa: F401
    from clvln_b200 import ops, utils, _lib
    from clvln_b200.agent import build_agent
    from clvln_b200.engine import TrainStep
    from clvln_b200.engine.graphs import GraphedTrainStep
    from clvln_b200.environ import R2RBatch
    import random
    _lib.lib()
    torch.backends.cuda.matmul.allow_tf32 = False
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))

    world, items = build_world(args.small, dev, lengths=args.lengths)
    cfg = utils.agent_cfg("ENVDROP")
    cfg.TRAIN.BATCH_SIZE = args.batch
    random.seed(2020)
    env = R2RBatch(world, items, batch_size=args.batch, device=dev, rank=rank, world_size=world_size)
    torch.manual_seed(2020)
    agent = build_agent(cfg, utils.StubTokenizer(), dev)
    agent.env = env
    agent.train()
    agent.sync_every = 0                                  # fixed-length sampled rollouts: no host polls
    step = GraphedTrainStep(cfg, agent) if args.graph else TrainStep(cfg, agent)
    if args.graph:
        step.prefetch_next = True                         # e2e: the next minibatch is assembled + copied under this step's kernels

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, staged, read_loss):
        if staged:
            env.prefetch(n)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0 = ops.CALLS[0]
        e0.record()
        for _ in range(n):
            loss = step()
            if read_loss:
                loss.item()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev)
        if world_size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t), ops.CALLS[0] - c0

    timed(max(3, args.warmup), False, False)              # warm-up (also captures the graphs)
    clk = ClockSampler(local)
    if rank == 0:
        clk.start()
    t_dev, launches = timed(args.steps, True, False)      # inputs resident in HBM before the region
    t_e2e, _ = timed(args.steps, False, True)             # public API: H2D per step + loss read-back
    clocks = clk.stop() if rank == 0 else None
    h2d = env._last_ib.h2d_bytes
    n_ep = args.batch * world_size * args.steps
    out = {
        "metric": METRIC, "value": round(n_ep / t_dev, 2), "unit": UNIT, "n_gpus": world_size, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": round(t_dev / args.steps * 1e3, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 state/accumulate; bf16 feature table; forward + input-gradient GEMMs bf16x3 on tcgen05 (3 bf16 MMAs per product, ~2e-5 rel); weight-gradient GEMMs and the encoder's embedding-side input gradient tcgen05 kind::tf32 (csrc/wgrad.cu, fp32 accumulate); no library GEMM in the iteration",
        "data": "synthetic",
        "config": {"workload": WORKLOAD % args.batch,
                   "global_batch": args.batch * world_size, "parallelism": f"dp{world_size}",
                   "table": "%d viewpoints x 36 x 2048 bf16 = %.2f GB in HBM" % (world.n_vp, world.n_vp * 36 * 2048 * 2 / 1e9),
                   "l2": "inputs larger than L2: every step gathers random viewpoints from the %.2f GB table" % (world.n_vp * 36 * 2048 * 2 / 1e9),
                   "cuda_graph": bool(args.graph),
                   "instruction_lengths": "80 for every row" if args.lengths == "80" else "real R2R training distribution (mean 31.3)"},
        "e2e": {"value": round(n_ep / t_e2e, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": round(t_e2e / args.steps * 1e3, 3)},
        "gpu_launches": launches, "clocks": clocks,
    }
    if rank == 0:
        store = agent.store_of(env)
        out["roofline"] = roofline_pano(store, ops, torch, args.batch, agent.split_for(args.batch), peaks)
        if args.batch == 64:
            out["roofline"]["traffic"] = ncu_traffic("r01_ncu_pano_cluster_B64.csv")
            out["roofline"]["traffic_source"] = "profiles/r01_ncu_pano_cluster_B64.csv (ncu --set full, bytes per launch; algorithmic = 64 x 147456 = 9437184 + 589824 B of keep-bits)"
        # the same kernel at larger episode counts per launch (BASELINE config 5): where it leaves the latency regime
        out["roofline"]["sweep"] = [
            {k: r[k] for k in ("episodes_per_launch", "us_per_launch", "achieved", "frac")}
            for r in (roofline_pano(store, ops, torch, b, 1, peaks) for b in (256, 1024, 2048))]
        if world_size == 1:
            out["roofline"]["others"] = roofline_others(ops, torch, args.batch, peaks)
        if world_size == 1 and not args.no_cpu_baseline:
            threads = host_threads()
            from oracle import ref_loader
            B = args.batch
            real = ref_loader.reference_available()
            if real:        # the unmodified reference (oracle/_ref) on this very world and episode set
                it = reference_iteration_factory(world, items, B, threads)
                desc = "%d-viewpoint world; the unmodified reference agent + trainer recipe from oracle/_ref" % world.n_vp
            else:
                from clvln_b200.environ import make_world, make_items
                w_small = make_world(n_scans=8, seed=2020)
                it = cpu_iteration_factory(w_small, make_items(w_small, 4 * B, seed=2020, fixed_len=80), B, threads)
                desc = "8-scan synthetic world; oracle CPU restatement of the reference"
            it()
            n, t0 = 0, time.time()
            while n < 2 or (time.time() - t0 < 15.0 and n < 50):
                it()
                n += 1
            dt = time.time() - t0
            out["cpu_baseline"] = {"value": round(B * n / dt, 3), "unit": UNIT, "cores": threads, "cpu": cpu_model(),
                                   "kind": "reference" if real else "port",
                                   "sample": f"{n} full EnvDrop training iterations at B={B}, L=80 ({desc}; fp32, {threads} torch threads, {dt:.1f} s)"}
        print(json.dumps(out), flush=True)
    if world_size > 1:
        # the iteration graphs hold captured NCCL kernels: release them before the communicator, and never let a stuck
        # teardown keep the job alive (the result line is out)
        sys.stdout.flush()
        threading.Timer(20.0, lambda: os._exit(0)).start()
        torch.cuda.synchronize()
        if hasattr(step, "graphs"):
            step.graphs.clear()
        del step
        dist.barrier()
        dist.destroy_process_group()
        os._exit(0)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
