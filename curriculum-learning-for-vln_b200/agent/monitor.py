"""Self-Monitoring agent (Ma et al., ICLR 2019) on the device-resident rollout.

Mirrors src/agent/monitor.py (ctor :25-66, _instr_variable :68-87 — always the full
``max_enc_len`` tokens because the progress head is Linear(max_enc_len + H, 1) —, rollout
:89-199): candidate-only decoder, action CE + progress-monitor MSE mixed by ``lamb``.
"""
import torch

from .. import ops
from ..model import EncoderLSTM, MonitorDecoder
from ..model.units import LengthMask
from .base import BaseAgent, RolloutState
from .follower import masked_mean_ce


class SelfMonitorAgent(BaseAgent):
    def __init__(self, model_cfg, max_enc_len, results_dir, device, env, tokenizer, episode_len=10):
        super().__init__(results_dir, device, env, tokenizer, episode_len=episode_len)
        self.cfg = model_cfg
        self.action_emb_size = self.feature_size
        self.max_enc_len = max_enc_len
        self.encoder = EncoderLSTM(tokenizer.vocab_size(), model_cfg.WORD_EMB_SIZE, model_cfg.HIDDEN_SIZE,
                                   padding_idx=0, drop_ratio=model_cfg.DROP_RATE,
                                   bidirectional=model_cfg.ENC_BIDIRECTION, num_layers=model_cfg.ENC_LAYERS)
        self.decoder = MonitorDecoder(rnn_hidden_size=model_cfg.HIDDEN_SIZE, drop_ratio=model_cfg.DROP_RATE,
                                      max_enc_len=max_enc_len, mlp_dims=list(model_cfg.MLP_HIDDEN),
                                      action_embed_size=self.action_emb_size, feature_size=self.feature_size)
        self.progress_losses = []
        self._finish_init()

    def _modules(self):
        return [self.encoder, self.decoder]

    def reset_loss(self):
        self.losses = []
        self.progress_losses = []

    # ---- beam search hooks (monitor.py:201-225) -----------------------------------------------------------------------
    beam_full_length = True            # the decoder's progress head needs the full-width instruction (rollout, above)

    def beam_start_state(self, h_t):
        return torch.zeros(h_t.shape[0], self.action_emb_size, device=self.device)

    def decode_observation(self, store, vp, view, h_t, c_t, a_prev, ctx, ctx_mask, ended=None):
        from .follower import _beam_next_action
        cands, lens = ops.gather_cand(store, vp, view)
        cmask = LengthMask(lens, ops.NSLOT)
        (logit, _), (h_t, c_t), _ = self.decoder(None, a_prev, cands, h_t, c_t, ctx, ctx_mask, cmask)
        logit = logit.masked_fill(cmask.dense(), float("-inf"))
        action = _beam_next_action(store, vp, logit, ended)
        return logit, h_t, c_t, ops.gather_action_feat(store, vp, view, action, None).detach()

    decode_obervation = decode_observation

    def rollout(self, train_ml=True, train_cl=False, reset=True, restart=False, lamb=0.5, speaker=None,
                avoid_cyclic=False, feedback="sample", return_traj=None):
        assert speaker is None and not avoid_cyclic, "speaker / avoid_cyclic paths are not part of this build"
        ib = self.env.reset_index(restart=restart, full_length=True)
        store = self.store_of(self.env)
        B = ib.vp.shape[0]
        T, poll = self._horizon(ib, feedback)
        ctx, h_t, c_t = self.encoder(ib.tokens, ib.lengths)
        ctx_mask = LengthMask(ib.lengths, ctx.shape[1])
        st = RolloutState(store, ib, T)
        training = self.encoder.training
        a_prev = torch.zeros(B, self.action_emb_size, device=self.device)
        start_dist = st.dist[0]
        ml, prog_log = 0.0, torch.zeros((), device=self.device)
        for t in range(T):
            cands, lens = ops.gather_cand(store, st.vp[t], st.view[t])
            # the reference pads candidates to the longest row of the batch (base.py:150-151) and its BatchNorm
            # statistics run over exactly those rows: the decoder takes that width from `lens` on the device
            # (masked statistics over all 16 slots), so the rollout never reads it back to the host
            C = ops.NSLOT
            cmask = LengthMask(lens, C)
            (logit, prog), (h_t, c_t), _ = self.decoder(None, a_prev, cands, h_t, c_t, ctx, ctx_mask, cmask)
            logit = logit.masked_fill(cmask.dense(), float("-inf"))
            target = st.teacher
            off = self.rng.next() if feedback == "sample" else 0
            ce, _, _, action = ops.policy_head(logit, target, feedback, self.rng, off)
            act_loss = ce if train_cl else masked_mean_ce(ce, target)
            if t == 0:
                cur = act_loss
            else:
                cur_dist, ended = st.dist[t], st.ended[t].bool()
                pt = (start_dist - cur_dist) / start_dist
                pt = torch.where(cur_dist <= 3.0, torch.ones_like(pt), pt)
                pt = torch.where(ended, prog.detach(), pt)
                pl = (prog - pt) ** 2
                if not train_cl:
                    pl = pl.mean()
                prog_log = prog_log + pl.detach().mean()
                cur = lamb * pl + (1 - lamb) * act_loss
            ml = ml + cur
            if self.trace is not None:
                self.trace.append(dict(logits=logit.detach(), target=target, action=action))
            a_prev = ops.gather_action_feat(store, st.vp[t], st.view[t], action, st.ended[t])
            st.step(t, action)
            if poll and (t + 1) % poll == 0 and t + 1 < T and st.all_ended(t):
                break
        self.ml_loss = ml
        self.progress_loss = prog_log
        if not train_cl and not restart:
            self.losses.append(ml.detach())
        if not restart:
            self.progress_losses.append(prog_log)
        self.last_state, self.last_batch = st, ib
        if return_traj if return_traj is not None else not training:
            return self._trajectories(st)
        return []
