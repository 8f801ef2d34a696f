"""Speaker-Follower agent (Fried et al., NeurIPS 2018) on the device-resident rollout.

Mirrors src/agent/follower.py (ctor :26-63, rollout :65-173): panoramic visual attention with
the previous action's feature as LSTM input, teacher / argmax / sample feedback, CE loss summed
over steps — mean over the running episodes per step (:62,:127), or the per-episode vector when
``train_cl`` (:63,:128).
"""
import torch

from .. import ops
from ..model import EncoderLSTM, AttnDecoderLSTM
from ..model.units import LengthMask
from .base import BaseAgent, RolloutState


def masked_mean_ce(ce, target):
    """nn.CrossEntropyLoss(ignore_index=-1) with mean reduction: sum over non-ignored / their count."""
    n = (target >= 0).sum().clamp(min=1)
    return ce.sum() / n


def _beam_next_action(store, vp, logit, ended):
    """follower.py:190-195: the argmax slot, -1 for STOP (slot n_cand) and for rows whose search has ended."""
    action = logit.argmax(1).to(torch.int32)
    stop = action == store.n_cand[vp.long()]
    if ended is not None:
        stop = stop | torch.as_tensor(ended, dtype=torch.bool, device=logit.device)
    return torch.where(stop, torch.full_like(action, -1), action)


class FollowerAgent(BaseAgent):
    def __init__(self, model_cfg, results_dir, device, env, tokenizer, glove=None, episode_len=10):
        super().__init__(results_dir, device, env, tokenizer, episode_len=episode_len)
        self.cfg = model_cfg
        self.action_emb_size = self.feature_size
        self.encoder = EncoderLSTM(tokenizer.vocab_size(), model_cfg.WORD_EMB_SIZE, model_cfg.HIDDEN_SIZE,
                                   padding_idx=0, drop_ratio=model_cfg.DROP_RATE,
                                   bidirectional=model_cfg.ENC_BIDIRECTION, num_layers=model_cfg.ENC_LAYERS,
                                   glove=glove)
        self.decoder = AttnDecoderLSTM(hidden_size=model_cfg.HIDDEN_SIZE, drop_ratio=model_cfg.DROP_RATE,
                                       action_embed_size=self.action_emb_size, feature_size=self.feature_size)
        self._finish_init()

    def _modules(self):
        return [self.encoder, self.decoder]

    # ---- beam search hooks (follower.py:175-198) ----------------------------------------------------------------------
    def beam_start_state(self, h_t):
        return torch.zeros(h_t.shape[0], self.action_emb_size, device=self.device)

    def decode_observation(self, store, vp, view, h_t, c_t, a_prev, ctx, ctx_mask, ended=None):
        pano = ops.PanoView(store, vp, view)
        pano.split = self.split_for(vp.shape[0])
        logit, (h_t, c_t), _ = self.decoder(pano, a_prev, ops.CandView(store, vp, view), h_t, c_t, ctx, ctx_mask)
        action = _beam_next_action(store, vp, logit, ended)
        return logit, h_t, c_t, ops.gather_action_feat(store, vp, view, action, None).detach()

    decode_obervation = decode_observation

    def rollout(self, train_ml=True, train_rl=False, train_cl=False, reset=True, restart=False, speaker=None,
                avoid_cyclic=False, feedback="sample", return_traj=None):
        assert speaker is None and not avoid_cyclic, "speaker / avoid_cyclic paths are not part of this build"
        ib = self.env.reset_index(restart=restart)
        store = self.store_of(self.env)
        B = ib.vp.shape[0]
        T, poll = self._horizon(ib, feedback)
        ctx, h_t, c_t = self.encoder(ib.tokens, ib.lengths)
        ctx_mask = LengthMask(ib.lengths, ctx.shape[1])
        st = RolloutState(store, ib, T)
        training = self.encoder.training
        a_prev = torch.zeros(B, self.action_emb_size, device=self.device)
        ml = torch.zeros(B, device=self.device) if train_cl else torch.zeros((), device=self.device)
        for t in range(T):
            pano = st.pano(t)
            pano.split = self.split_for(st.B)
            logit, (h_t, c_t), _ = self.decoder(pano, a_prev, st.cands(t), h_t, c_t, ctx, ctx_mask)
            target = st.teacher
            off = self.rng.next() if feedback == "sample" else 0
            ce, _, _, action = ops.policy_head(logit, target, feedback, self.rng, off)
            ml = ml + (ce if train_cl else masked_mean_ce(ce, target))
            if self.trace is not None:
                self.trace.append(dict(logits=logit.detach(), target=target, action=action))
            a_prev = ops.gather_action_feat(store, st.vp[t], st.view[t], action, st.ended[t])
            st.step(t, action)
            if poll and (t + 1) % poll == 0 and t + 1 < T and st.all_ended(t):
                break
        self.ml_loss = ml
        if not train_cl and not restart:
            self.losses.append(ml.detach())
        self.last_state, self.last_batch = st, ib
        if return_traj if return_traj is not None else not training:
            return self._trajectories(st)
        return []
