"""Agent base: result bookkeeping, the test() loop and the device-side rollout scaffolding.

Public protocol = the reference's (src/agent/base.py:22-112): ``env`` (assigned by the trainer),
``results``, ``losses``, ``rollout(**kw)``, ``test(iters=None, **kw)``, ``get_results()``,
``write_results()``, ``train()/eval()``, ``save_model/load_model``, ``trainable_params``.

A rollout here never leaves the device: the episode state (viewpoint, view, ended, distance)
lives in [T+1,B] trajectory buffers advanced by the vln_env_step kernel, the observation
"tensors" of base.py:114-178 are views of the HBM feature table, and the only host round trip is
an optional `all ended?` poll every ``sync_every`` steps (the reference polls every step through
``.cpu()``, envdrop.py:198,219).  Teacher-forced rollouts need no poll at all — their length is
known on the host (longest shortest-path in the batch + the STOP step).
"""
import json
import os
import random

import torch

from .. import ops
from ..model import units as U


class RolloutState:
    """[T+1,B] device trajectory: row t is the state before step t."""

    def __init__(self, store, ib, T):
        B, dev = ib.vp.shape[0], ib.vp.device
        self.store, self.goal, self.B, self.T = store, ib.goal, B, T
        # zero-filled: rows that stop taking part in a rollout early (paired rollouts) must stay valid table indices
        self.vp = torch.zeros((T + 1, B), dtype=torch.int32, device=dev)
        self.view = torch.zeros((T + 1, B), dtype=torch.int32, device=dev)
        self.ended = torch.zeros((T + 1, B), dtype=torch.uint8, device=dev)
        self.dist = torch.zeros((T + 1, B), dtype=torch.float32, device=dev)
        self.n_active = torch.zeros((T,), dtype=torch.int32, device=dev)
        self.vp[0].copy_(ib.vp)
        self.view[0].copy_(ib.view)
        self.teacher, d0 = ops.env_observe(store, self.vp[0], self.ended[0], self.goal)
        self.dist[0].copy_(d0)
        self.steps = 0

    def pano(self, t):
        return ops.PanoView(self.store, self.vp[t], self.view[t])

    def cands(self, t):
        return ops.CandView(self.store, self.vp[t], self.view[t])

    def step(self, t, action):
        """Advance to row t+1; returns (reward, mask) of the transition and updates ``teacher``."""
        _, _, _, _, self.teacher, reward, mask = ops.env_step(
            self.store, self.vp[t], self.view[t], self.ended[t], self.dist[t], self.goal, action,
            out=(self.vp[t + 1], self.view[t + 1], self.ended[t + 1], self.dist[t + 1]),
            n_active=self.n_active[t:t + 1])
        self.steps = t + 1
        return reward, mask

    def all_ended(self, t):
        """Host poll (one 4-byte D2H sync): did every episode end by step t?"""
        return int(self.n_active[t].item()) == 0


class BaseAgent:
    ignore_id = -1

    def __init__(self, results_dir, device, env, tokenizer, img_feat_size=2048, angle_feat_size=128,
                 episode_len=20):
        self.env = env
        self.results_save_dir = results_dir
        random.seed(1)                                   # base.py:28 — part of the ordering contract
        self.results = {}
        self.losses = []
        self.device = torch.device(device) if not isinstance(device, torch.device) else device
        self.tokenizer = tokenizer
        self.episode_len = episode_len
        self.img_feat_size, self.angle_feat_size = img_feat_size, angle_feat_size
        self.feature_size = img_feat_size + angle_feat_size
        self.rng = None
        self._stores = {}
        self.sync_every = 4            # poll `all ended` every k steps in student-forced rollouts (0 = never)
        self.fixed_steps = None        # run exactly this many steps (CUDA-graph capture), no polling
        self.pano_split = None         # panorama kernel variant (vln_pano_attn `split`); None = automatic
        self.last_state = None
        self.trace = None              # set to a list to record per-step logits / targets / actions

    # ---- plumbing ----------------------------------------------------------------------------
    def _modules(self):
        raise NotImplementedError

    def _finish_init(self):
        for m in self._modules():
            m.to(self.device)
        if self.device.type == "cuda":
            # data parallel: every replica draws its own dropout masks / action samples (rank folded into the Philox
            # seed); rank 0 keeps the single-GPU stream
            import torch.distributed as dist
            rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
            self.rng = ops.Rng(2020 + 7919 * rank, self.device)
            for m in self._modules():
                U.use_rng(m, self.rng)

    def store_of(self, env):
        """FeatureStore (HBM tables) of the env's world, created once per world."""
        key = id(env.world)
        if key not in self._stores:
            self._stores[key] = ops.FeatureStore.from_world(env.world, self.device)
        return self._stores[key]

    def split_for(self, B):
        """Kernel variant of vln_pano_attn (1 = chosen by the library from B, 2 = cluster, 4 = streaming)."""
        return 1 if self.pano_split is None else self.pano_split

    def _horizon(self, ib, feedback):
        if self.fixed_steps is not None:
            return self.fixed_steps, 0
        if feedback == "teacher":
            return min(self.episode_len, ib.teacher_steps), 0
        return self.episode_len, self.sync_every

    def _trajectories(self, st):
        n = st.steps + 1
        return self.env.traj_from_index(st.vp[:n].cpu().numpy(), st.view[:n].cpu().numpy())

    # ---- reference protocol ----------------------------------------------------------------------
    def write_results(self, split="train"):
        path = os.path.join(self.results_save_dir, "%s.json" % split)
        with open(path, "w") as f:
            json.dump(self.get_results(), f)

    def get_results(self):
        return [{"instr_id": k, "trajectory": v} for k, v in self.results.items()]

    def rollout(self, **kw):
        raise NotImplementedError

    # ---- beam search (base.py:183-482; agent/beam.py) ---------------------------------------------------------------
    def running_state(self, h_t, c_t, extra=None, **kw):
        """What a search state carries between expansions (the agents' overrides, e.g. envdrop.py:280-281)."""
        if extra is None:
            extra = kw.get("h_tilde", kw.get("a_t_prev"))
        return (h_t, c_t, extra)

    def beam_start_state(self, h_t):
        """The third element of the start states' running state ([B, .])."""
        raise NotImplementedError

    def decode_observation(self, store, vp, view, h_t, c_t, extra, ctx, ctx_mask, ended):
        """One decoder step for a batch of (viewpoint, view) states -> (masked logits [B, 16], h_t, c_t, extra)."""
        raise NotImplementedError

    decode_obervation = decode_observation                   # (the reference's spelling, base.py:472)

    def _dijkstra(self, max_candidates):
        from . import beam
        return beam.dijkstra(self, max_candidates)

    def beam_rollout(self, speaker, beam_size):
        from . import beam
        return beam.beam_rollout(self, speaker, beam_size)

    def beam_search(self, speaker, beam_size=30):
        from . import beam
        return beam.beam_search(self, speaker, beam_size)

    def test(self, iters=None, **kw):
        """base.py:63-82: roll out until an instr_id repeats (or `iters` batches)."""
        self.env.reset_epoch(shuffle=(iters is not None))
        self.losses = []
        self.results = {}
        kw.setdefault("return_traj", True)
        if iters is not None:
            for _ in range(iters):
                for traj in self.rollout(**kw):
                    self.results[traj["instr_id"]] = traj["path"]
            return
        looped = False
        while not looped:
            for traj in self.rollout(**kw):
                if traj["instr_id"] in self.results:
                    looped = True
                else:
                    self.results[traj["instr_id"]] = traj["path"]

    def train(self):
        for m in self._modules():
            m.train()

    def eval(self):
        for m in self._modules():
            m.eval()

    def reset_loss(self):
        self.losses = []

    def trainable_params(self):
        out = []
        for m in self._modules():
            out += [p for p in m.parameters() if p.requires_grad]
        return out

    def _ckpt_names(self):
        return ["encoder", "decoder"]

    def save_model(self, path, **extra):
        out = dict(extra)
        for n in self._ckpt_names():
            out[f"{n}_state_dict"] = getattr(self, n).state_dict()
        torch.save(out, path)

    def load_model(self, path, ret=True, cuda=0):
        ckpt = torch.load(path, map_location=self.device, weights_only=False)
        for n in self._ckpt_names():
            getattr(self, n).load_state_dict(ckpt[f"{n}_state_dict"])
        if ret:
            return ckpt
