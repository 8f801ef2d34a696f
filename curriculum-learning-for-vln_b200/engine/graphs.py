"""CUDA-graph replay of a training iteration.

The reference's iteration is host-driven (one `.cpu()` per decoder step, envdrop.py:198); on a
B200 the ~2 000 small launches of an EnvDrop iteration (two rollouts + backward) are bound by the
Python/launch path, not by the GPU.  ``GraphedTrainStep`` captures zero_grad -> rollouts ->
backward once per teacher-rollout length into a CUDA graph and replays it: the minibatch's index
tensors are copied into static buffers, the Philox base lives in device memory and is advanced
inside the graph, so every replay sees new data and new dropout masks / samples.  The gradient
all-reduce (by default, see ``dp_in_graph``) and the fused clip + update stay outside the graph (3 launches).
"""
import torch

from .. import ops
from .trainer import TrainStep


class _StaticEnv:
    """Stands in for agent.env while capturing/replaying: hands out the static IndexBatch."""

    def __init__(self, env, ib):
        self._env, self._ib = env, ib

    def reset_index(self, **kw):
        return self._ib

    def __getattr__(self, k):
        return getattr(self._env, k)


class GraphedTrainStep(TrainStep):
    def __init__(self, cfg, agent, optimizer=None, weights=None, warmup=2):
        super().__init__(cfg, agent, optimizer, weights)
        self.graphs = {}
        self.static = None
        self.warmup = warmup
        self.full_length = cfg.MODEL.NAME == "SELF-MONITOR"
        # draw + stage the NEXT minibatch right after this iteration is launched, so that the host-side assembly
        # and its H2D copies overlap the GPU's work instead of following the caller's loss read-back.  The order of
        # minibatches is unchanged; callers switch it off for the last iteration before they touch the env / the
        # global `random` stream themselves (epoch end: evaluation, curriculum round switch).
        self.prefetch_next = False
        # Data parallel.  Default: the gradient all-reduce stays OUTSIDE the captured graph — `opt.step()` issues exactly
        # one all-reduce of the flat buffer per iteration on every rank, whatever mix of warm-up / capture / replay the
        # ranks are in (the round-1 scheme, measured at 8 GPUs).  VLN_DP_INGRAPH=1 captures the bucketed, overlapped
        # all-reduce into the graph instead (decoder + critic bucket under the encoder's backward): then warm-up
        # iterations all-reduce for real and capture iterations do not, so it is only correct when every rank meets new
        # graph keys at the same iterations — which environ/batch.py guarantees by taking the teacher-rollout length over
        # the GLOBAL minibatch.  (Before that guarantee a 4-GPU run hung: one rank had warmed up a new length while its
        # peers replayed, and the collectives no longer paired up.)  Measured at 2 GPUs the in-graph variant is not faster
        # (4.62 vs 4.49 ms single; round 1 outside the graph: 4.89 vs 4.78), so it is opt-in.
        import os
        self.dp_in_graph = os.environ.get("VLN_DP_INGRAPH", "0") == "1"
        fd = getattr(agent, "_fused", None)
        if fd is not None and not self.dp_in_graph:
            fd.on_grads_ready = None        # (TrainStep set the early-bucket hook; it must not be captured)
        # warm-up iterations run on a side stream and the capture on its own: the AccumulateGrad nodes of the (persistent)
        # parameters then see a different stream than the one they were created on; intended here, so silence the notice
        fn = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
        if fn is not None:
            fn(False)

    def _body(self):
        self.opt.zero_grad()
        loss, item = self.losses()
        loss.backward()
        if self.dp_in_graph:
            self.opt.finish_reduce()        # data parallel, opt-in: the rest of the gradient all-reduce, inside the graph
        self.agent.rng.advance()
        return loss.detach(), item

    def _load(self, ib):
        import copy
        if self.static is None:
            L = self.agent.env.max_len
            st = copy.copy(ib)
            st.tokens = torch.zeros((ib.tokens.shape[0], L), dtype=torch.int64, device=ib.tokens.device)
            for k in ("lengths", "vp", "view", "goal", "index"):
                setattr(st, k, getattr(ib, k).clone())
            self.static = st
        st = self.static
        st.tokens.zero_()
        st.tokens[:, :ib.tokens.shape[1]].copy_(ib.tokens)
        for k in ("lengths", "vp", "view", "goal", "index"):
            getattr(st, k).copy_(getattr(ib, k))
        st.teacher_steps = ib.teacher_steps
        return st

    def __call__(self):
        ag = self.agent
        env = ag.env
        ib = env.reset_index(full_length=self.full_length)
        st = self._load(ib)
        key = (min(ag.episode_len, ib.teacher_steps), id(env.world))
        ag.env = _StaticEnv(env, st)
        saved = ag.sync_every
        ag.sync_every = 0
        try:
            if key not in self.graphs:
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    for _ in range(self.warmup):
                        ag.rng.off = 0
                        self._body()
                torch.cuda.current_stream().wait_stream(s)
                g = torch.cuda.CUDAGraph()
                ag.rng.off = 0
                # the optimiser updates parameters through raw pointers between replays: make the capture
                # re-derive every cached bf16 weight split, so the split kernels are part of the graph
                ops.WEIGHT_EPOCH[0] += 1
                c0 = ops.CALLS[0]
                logs = getattr(ag, "logs", None)
                n_log = {k: len(v) for k, v in logs.items()} if logs is not None else {}
                with torch.cuda.graph(g):
                    loss, item = self._body()
                # device scalars the rollout appended to agent.logs (entropy / critic_loss / total, envdrop.py:193,257,
                # 262): static outputs of the graph — every replay appends a copy, the capture's own entries go
                stat = {}
                if logs is not None:
                    for k, v in logs.items():
                        new = v[n_log.get(k, 0):]
                        if new and all(torch.is_tensor(x) for x in new):
                            stat[k] = list(new)
                        del v[n_log.get(k, 0):]
                # the rollout state / batch objects of THIS graph (static buffers the replay rewrites)
                refs = (getattr(ag, "last_state", None), getattr(ag, "last_batch", None),
                        getattr(getattr(ag, "_fused", None), "last", None))
                self.graphs[key] = (g, loss, item, ops.CALLS[0] - c0, stat, refs)
            g, loss, item, n_calls, stat, refs = self.graphs[key]
            g.replay()
            ops.CALLS[0] += n_calls
            if self.opt.world > 1:
                if self.dp_in_graph:
                    self.opt._reduced = [(0, self.opt.grad.numel())]  # the replayed graph all-reduced the whole buffer
                else:
                    self.opt._pending, self.opt._reduced = [], []     # nothing reduced yet: opt.step() all-reduces once
            ag.last_state, ag.last_batch = refs[0], refs[1]
            if getattr(ag, "_fused", None) is not None:
                ag._fused.last = refs[2]
            for k, refs in stat.items():
                ag.logs[k] += [x.clone() for x in refs]
        finally:
            ag.env = env
            ag.sync_every = saved
        self.opt.step()
        if item is not None:
            self.weights.record(st.index, item)
        if self.prefetch_next and not env._staged:
            env.prefetch(1, full_length=self.full_length)
        # `loss` is the graph's static output buffer: hand out a copy, so that values kept across iterations (the
        # trainers stack an epoch's losses and read them back once) are not overwritten by the next replay
        return loss.clone()
