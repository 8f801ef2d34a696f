"""EnvDrop agent (Tan et al., NAACL 2019): imitation (CE) + A2C, on the device-resident rollout.

Mirrors src/agent/envdrop.py (ctor :27-73, rollout :86-278, save/load :298-313) — same
arguments, same ``self.loss`` / ``self.ml_loss`` / ``self.rl_loss`` / ``self.logs`` outputs —
with the per-step host work (numpy feature assembly :75-84, H2D copies, `.cpu()` of the actions
:198, numpy reward shaping :207-219, the per-step critic calls of the A2C loop :240-264) replaced
by kernels and one batched critic call.  Dead reference paths (speaker back-translation :104-120,
avoid_cyclic :167-172) are not carried over.
"""
import dataclasses
from collections import defaultdict

import torch

from .. import ops
from ..model import EncoderLSTM, EnvDropDecoder, Critic
from ..model.units import LengthMask
from .base import BaseAgent, RolloutState
from .fused import FusedDecoder


class EnvDropAgent(BaseAgent):
    def __init__(self, model_cfg, max_enc_len, results_dir, device, env, tokenizer, episode_len=20):
        super().__init__(results_dir, device, env, tokenizer, episode_len=episode_len)
        self.cfg = model_cfg
        self.action_emb_size = model_cfg.ACT_EMB_SIZE
        self.max_enc_len = max_enc_len
        self.encoder = EncoderLSTM(tokenizer.vocab_size(), model_cfg.WORD_EMB_SIZE, model_cfg.HIDDEN_SIZE,
                                   padding_idx=0, drop_ratio=model_cfg.DROP_RATE,
                                   bidirectional=model_cfg.ENC_BIDIRECTION, num_layers=model_cfg.ENC_LAYERS)
        self.decoder = EnvDropDecoder(hidden_size=model_cfg.HIDDEN_SIZE, drop_ratio=model_cfg.DROP_RATE,
                                      feat_drop_ratio=model_cfg.FEAT_DROP_RATE, action_embed_size=self.action_emb_size,
                                      angle_feat_size=self.angle_feat_size, feature_size=self.feature_size)
        self.critic = Critic(hidden_size=model_cfg.HIDDEN_SIZE, drop_ratio=model_cfg.DROP_RATE)
        self.logs = defaultdict(list)
        self.loss = {}
        self.fused = True              # decoder rollout as one hand-differentiated node (agent/fused.py)
        self._fused = FusedDecoder(self.decoder)
        self._finish_init()

    def _modules(self):
        return [self.encoder, self.decoder, self.critic]

    def _ckpt_names(self):
        return ["encoder", "decoder", "critic"]

    def reset_loss(self):
        self.losses = []
        self.logs = defaultdict(list)

    def _decode(self, st, t, h_tilde, h_t, c_t, ctx, ctx_mask):
        pano, cands = st.pano(t), st.cands(t)
        pano.split = self.split_for(st.B)
        pose = ops.pose_feature(st.store, st.view[t])
        logit, (h_t, c_t), h_tilde = self.decoder(pose, pano, cands, h_tilde, h_t, c_t, ctx, ctx_mask)
        return logit, h_t, c_t, h_tilde

    # ---- beam search hooks (envdrop.py:280-297) ------------------------------------------------------------------------
    def beam_start_state(self, h_t):
        return h_t                                            # h_tilde starts as the encoder's h_t (base.py:236)

    def decode_observation(self, store, vp, view, h_t, c_t, h_tilde, ctx, ctx_mask, ended=None):
        pano, cands = ops.PanoView(store, vp, view), ops.CandView(store, vp, view)
        pano.split = self.split_for(vp.shape[0])
        logit, (h_t, c_t), h_tilde = self.decoder(ops.pose_feature(store, view), pano, cands, h_tilde, h_t, c_t, ctx, ctx_mask)
        return logit, h_t, c_t, h_tilde

    decode_obervation = decode_observation

    def rollout_pair(self, train_cl=False):
        """The two rollouts of one EnvDrop training iteration (trainer.py:411-421: teacher-forced for the imitation
        loss, then sampled on the SAME minibatch for A2C) stepped as ONE batch of 2B episodes: rows [0,B) sample,
        rows [B,2B) follow the teacher and only take part in the first T_teacher steps.  Same losses as two
        ``rollout`` calls (each half has its own dropout masks, as the two passes of the reference do), but the
        encoder and the first T_teacher decoder steps stream the weights once instead of twice.  Sets ``loss`` /
        ``ml_loss`` / ``rl_loss`` / ``logs`` like the two calls would."""
        assert self.fused and self.device.type == "cuda"
        ib = self.env.reset_index(restart=False)
        store = self.store_of(self.env)
        B = ib.vp.shape[0]
        T_t, _ = self._horizon(ib, "teacher")
        T, _ = self._horizon(ib, "sample")
        T_t = min(T_t, T)
        two = lambda x: torch.cat((x, x), 0)
        ib2 = dataclasses.replace(ib, tokens=two(ib.tokens), lengths=two(ib.lengths), lengths_cpu=two(ib.lengths_cpu),
                                  vp=two(ib.vp), view=two(ib.view), goal=two(ib.goal), index=two(ib.index))
        prep = self._fused.prepare(self.rng, 2 * B, T, "sample", True, self.device, pair=(B, T_t), L=ib2.tokens.shape[1])
        torch.cuda.nvtx.range_push("vln/encoder")
        ctx, h_t, c_t = self.encoder(ib2.tokens, ib2.lengths)
        torch.cuda.nvtx.range_pop()
        st = RolloutState(store, ib2, T + 1)
        torch.cuda.nvtx.range_push("vln/decoder_rollout")
        (ce, logps, ents, hiddens, logits, actions, targets, rewards, masks, last_h) = self._fused.run(
            self.rng, st, ctx, ib2.lengths, h_t, c_t, T, "sample", True, 0, self.split_for(2 * B), prep, pair=(B, T_t))
        torch.cuda.nvtx.range_pop()
        n = st.steps
        ml = ce[:T_t, B:].sum(0) if train_cl else ce[:T_t, B:].sum()
        if self.trace is not None:
            self.trace += [dict(logits=logits[t, B:], target=targets[t, B:], action=actions[t, B:]) for t in range(T_t)]
            self.trace += [dict(logits=logits[t, :B], target=targets[t, :B], action=actions[t, :B]) for t in range(n)]
        self.logs["entropy"].append(ents[:, :B].sum().detach())
        with torch.no_grad():
            last_value = self.critic(last_h[:B].contiguous())
        values = self.critic(hiddens[:, :B].reshape(n * B, -1)).view(n, B)
        loss_b, stats = ops.a2c_loss(logps[:, :B].contiguous(), ents[:, :B].contiguous(), values,
                                     rewards[:, :B].contiguous(), masks[:, :B].contiguous(), last_value,
                                     st.ended[n][:B].contiguous(), self.cfg.GAMMA, 0.01)
        rl = loss_b if train_cl else loss_b.sum()
        self.logs["total"].append(stats[0])
        self.logs["critic_loss"].append(stats[1])
        if self.cfg.RL_NORMALIZE == "total":
            rl = rl / stats[0]
        elif self.cfg.RL_NORMALIZE == "batch":
            rl = rl / B
        else:
            assert self.cfg.RL_NORMALIZE == "none"
        self.ml_loss, self.rl_loss = ml, rl
        self.loss = {"ml_loss": ml * self.cfg.ML_WEIGHT / B, "rl_loss": rl}
        if train_cl:
            self.losses.append(self.loss["ml_loss"].sum().detach())
        self.last_state = st
        self.last_batch = ib
        return []

    def rollout(self, train_ml=True, train_rl=False, train_cl=False, reset=True, restart=False, speaker=None,
                avoid_cyclic=False, feedback="sample", return_traj=None):
        assert speaker is None and not avoid_cyclic, "speaker / avoid_cyclic paths are not part of this build"
        if feedback != "sample":
            train_rl = False
        ib = self.env.reset_index(restart=restart)
        store = self.store_of(self.env)
        B = ib.vp.shape[0]
        T, poll = self._horizon(ib, feedback)
        use_fused = self.fused and self.device.type == "cuda"
        prep = None
        if use_fused:            # decoder stream offsets + the rollout's feature-dropout bits (side stream, overlaps the encoder)
            prep = self._fused.prepare(self.rng, B, T, feedback, train_rl, self.device, L=ib.tokens.shape[1])
        ctx, h_t, c_t = self.encoder(ib.tokens, ib.lengths)
        ctx_mask = LengthMask(ib.lengths, ctx.shape[1])
        st = RolloutState(store, ib, T + (1 if train_rl else 0))
        training = self.encoder.training

        if use_fused:
            (ce, logps, ents, hiddens, logits, actions, targets, rewards, masks, last_h) = self._fused.run(
                self.rng, st, ctx, ctx_mask.lengths, h_t, c_t, T, feedback, train_rl, poll, self.split_for(B), prep)
            n = st.steps
            ml = ce.sum(0) if train_cl else ce.sum()
            if self.trace is not None:
                self.trace += [dict(logits=logits[t], target=targets[t], action=actions[t]) for t in range(n)]
            if feedback == "sample":
                self.logs["entropy"].append(ents.sum().detach())
            if train_rl:
                with torch.no_grad():
                    last_value = self.critic(last_h)
                values = self.critic(hiddens.reshape(n * B, -1)).view(n, B)
        else:
            ml = torch.zeros(B, device=self.device) if train_cl else torch.zeros((), device=self.device)
            rewards, masks, hiddens, logps, ents = [], [], [], [], []
            h_tilde = h_t
            for t in range(T):
                logit, h_t, c_t, h_tilde = self._decode(st, t, h_tilde, h_t, c_t, ctx, ctx_mask)
                hiddens.append(h_t)
                off = self.rng.next() if feedback == "sample" else 0
                ce, logp, ent, action = ops.policy_head(logit, st.teacher, feedback, self.rng, off)
                ml = ml + (ce if train_cl else ce.sum())
                if self.trace is not None:
                    self.trace.append(dict(logits=logit.detach(), target=st.teacher, action=action))
                reward, mask = st.step(t, action)
                rewards.append(reward), masks.append(mask), logps.append(logp), ents.append(ent)
                if feedback == "sample":
                    self.logs["entropy"].append(ent.sum().detach())
                if poll and (t + 1) % poll == 0 and t + 1 < T and st.all_ended(t):
                    break
            n = st.steps
            if train_rl:
                with torch.no_grad():                               # envdrop.py:225-237: bootstrap value
                    _, last_h, _, _ = self._decode(st, n, h_tilde, h_t, c_t, ctx, ctx_mask)
                    last_value = self.critic(last_h)
                values = self.critic(torch.stack(hiddens).view(n * B, -1)).view(n, B)
                logps, ents, rewards, masks = (torch.stack(x) for x in (logps, ents, rewards, masks))
        self.ml_loss = ml

        rl = 0.0
        if train_rl:
            loss_b, stats = ops.a2c_loss(logps, ents, values, rewards, masks, last_value, st.ended[n],
                                         self.cfg.GAMMA, 0.01)
            rl = loss_b if train_cl else loss_b.sum()
            self.logs["total"].append(stats[0])
            self.logs["critic_loss"].append(stats[1])
            if self.cfg.RL_NORMALIZE == "total":
                rl = rl / stats[0]
            elif self.cfg.RL_NORMALIZE == "batch":
                rl = rl / B
            else:
                assert self.cfg.RL_NORMALIZE == "none"
        self.rl_loss = rl

        self.loss = {"ml_loss": ml * self.cfg.ML_WEIGHT / B if train_ml else 0.0,
                     "rl_loss": rl if train_rl else 0.0}
        if train_cl and not restart:
            val = self.loss["ml_loss"] + self.loss["rl_loss"]
            self.losses.append(0.0 if isinstance(val, float) else val.sum().detach())
        self.last_state = st
        self.last_batch = ib
        if return_traj if return_traj is not None else not training:
            return self._trajectories(st)
        return []
