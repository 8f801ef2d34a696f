"""Speaker: the instruction generator used for back-translation and for beam-search rescoring.

Public protocol = the reference's ``Speaker`` (src/agent/speaker.py:16-421): ``train(iters)``, ``get_insts()``, ``valid()``,
``from_shortest_path()``, ``teacher_forcing(train, features, insts, for_listener)``, ``infer_batch(sampling, train,
featdropmask)``, ``save(epoch, path)`` / ``load(path)`` with the reference's checkpoint layout, ``encoder`` / ``decoder``
modules with the reference's state_dict keys.

What differs is where the path lives: the reference walks the simulator along the shortest path and copies a
[B, T, 36, 2176] panorama tensor per batch to the device (speaker.py:191-226).  Here the walk is the device-resident
environment (vln_env_step with the teacher's action, T = longest shortest path + the STOP step, known on the host, so the
loop needs no read-back) and the panoramas stay in the HBM feature table: the encoder's attention reads them through
the fused gather + attention kernel (ops.PanoView), the action features through vln_gather_action_feat.
"""
import os

import numpy as np
import torch
import torch.nn.functional as F

from .. import ops
from ..model import units as U
from ..model.speaker import SpeakerDecoder, SpeakerEncoder
from .base import RolloutState


class Speaker:
    def __init__(self, spk_cfg, device, tok, env=None, img_feature_size=2048, angle_feat_size=128):
        self.device = torch.device(device) if not isinstance(device, torch.device) else device
        self.env = env
        self.img_feature_size, self.angle_feat_size = img_feature_size, angle_feat_size
        self.feature_size = img_feature_size + angle_feat_size
        self.tok = tok
        self.cfg = spk_cfg
        self.pad, self.bos = tok.word_to_index["<PAD>"], tok.word_to_index["<BOS>"]
        self.eos, self.unk = tok.word_to_index["<EOS>"], tok.word_to_index["<UNK>"]
        self.encoder = SpeakerEncoder(self.feature_size, spk_cfg.RNN_DIM, spk_cfg.DROPOUT, spk_cfg.BI_DIRECTION,
                                      angle_feat_size, spk_cfg.FEAT_DROPOUT).to(self.device)
        self.decoder = SpeakerDecoder(tok.vocab_size(), spk_cfg.WEMB, self.pad, spk_cfg.RNN_DIM,
                                      spk_cfg.DROPOUT).to(self.device)
        self.rng = ops.Rng(2021, self.device) if self.device.type == "cuda" else None
        for m in (self.encoder, self.decoder):
            U.use_rng(m, self.rng)
        self.encoder_optimizer = torch.optim.Adam(self.encoder.parameters(), lr=spk_cfg.LR)
        self.decoder_optimizer = torch.optim.Adam(self.decoder.parameters(), lr=spk_cfg.LR)
        self._stores = {}
        self._ib = None                      # the current minibatch (index form)
        self.last_path = None                # RolloutState of the last from_shortest_path (tests read it)

    # ---- plumbing --------------------------------------------------------------------------------------------------
    def store_of(self, env):
        key = id(env.world)
        if key not in self._stores:
            self._stores[key] = ops.FeatureStore.from_world(env.world, self.device)
        return self._stores[key]

    def _batch(self):
        if self._ib is None:
            self._ib = self.env.reset_index(full_length=True)
        return self._ib

    def reset(self, **kw):
        """Next minibatch of the env (the reference's ``self.env.reset()`` before every speaker call)."""
        self._ib = self.env.reset_index(full_length=True, **kw)
        return self._ib

    def _mode(self, train):
        for m in (self.encoder, self.decoder):
            m.train(bool(train))
        if train and self.rng is not None:
            self.rng.begin_iteration()

    # ---- speaker.py:75-88 --------------------------------------------------------------------------------------------
    def train(self, iters):
        for _ in range(iters):
            self.reset()
            self.encoder_optimizer.zero_grad()
            self.decoder_optimizer.zero_grad()
            loss = self.teacher_forcing(train=True)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(self.encoder.parameters(), 40.)
            torch.nn.utils.clip_grad_norm_(self.decoder.parameters(), 40.)
            self.encoder_optimizer.step()
            self.decoder_optimizer.step()

    # ---- speaker.py:90-123 -------------------------------------------------------------------------------------------
    def get_insts(self, wrapper=(lambda x: x)):
        self.env.reset_epoch(shuffle=True)
        path2inst = {}
        total = self.env.size()
        for _ in wrapper(range(total // self.env.batch_size + 1)):
            self.reset()
            insts = self.infer_batch()
            for item, inst in zip(self.env.batch, insts):
                if item["path_id"] not in path2inst:
                    path2inst[item["path_id"]] = self.shrink(inst)
        return path2inst

    def shrink(self, inst):
        """Tokenizer.shrink (misc.py:170-184): drop a leading <BOS> and everything from the first <EOS>."""
        inst = list(inst)
        if len(inst) == 0:
            return inst
        end = int(np.argmax(np.array(inst) == self.eos))
        start = 1 if len(inst) > 1 and inst[0] == self.bos else 0
        return inst[start:end]

    def valid(self, *aargs, **kwargs):
        path2inst = self.get_insts(*aargs, **kwargs)
        self.env.reset_epoch(shuffle=True)
        N = 1 if self.cfg.FAST_TRAIN else 3
        metrics = np.zeros(3)
        for _ in range(N):
            self.reset()
            metrics += np.array(self.teacher_forcing(train=False))
        metrics /= N
        return (path2inst, *metrics)

    # ---- speaker.py:191-226 ------------------------------------------------------------------------------------------
    def from_shortest_path(self, viewpoints=None, get_first_feat=False):
        """Walks the batch along its shortest paths on the device.  -> ((panoramas, can_feats[, first_feat]), lengths):
        panoramas = ops.PanoView over the B*T steps (row b*T + t), can_feats [B, T, 2176] = feature of the candidate the
        teacher picks (zeros for STOP / after the end), lengths int32 [B] = steps up to and including the STOP."""
        ib = self._batch()
        store = self.store_of(self.env)
        T = int(ib.teacher_steps)
        st = RolloutState(store, ib, T)
        B = st.B
        can = []
        for t in range(T):
            act = st.teacher                                # -1 after the end, n_cand[vp] = STOP, else the slot to take
            moves = (act >= 0) & (act < store.n_cand[st.vp[t].long()])
            feat = ops.gather_action_feat(store, st.vp[t], st.view[t], torch.clamp(act, min=0), None)
            can.append(feat * moves.unsqueeze(1).to(feat.dtype))
            st.step(t, act)
        self.last_path = st
        lengths = (T - st.ended[:T].to(torch.int32).sum(0)).to(torch.int32)
        pano = ops.PanoView(store, st.vp[:T].t().contiguous().view(-1), st.view[:T].t().contiguous().view(-1))
        can_feats = torch.stack(can, 1).contiguous()
        if viewpoints is not None:
            vps = st.vp[:T].cpu().numpy()
            for i in range(B):
                viewpoints[i].extend(self.env._vp_name(int(g)) for g in vps[:, i])
        if get_first_feat:
            first = torch.zeros((B, self.feature_size), device=self.device)
            first[:, -self.angle_feat_size:] = ops.pose_feature(store, ib.view)
            return (pano, can_feats, first), lengths
        return (pano, can_feats), lengths

    def gt_words(self, obs=None):
        return self._batch().tokens

    # ---- speaker.py:235-290 ------------------------------------------------------------------------------------------
    def teacher_forcing(self, train=True, features=None, insts=None, for_listener=False):
        self._mode(train)
        if features is not None:
            assert insts is not None
            (img_feats, can_feats), lengths = features
        else:
            (img_feats, can_feats), lengths = self.from_shortest_path()
        lengths = torch.as_tensor(lengths, dtype=torch.int32, device=self.device)
        batch_size = can_feats.shape[0]
        ctx = self.encoder(can_feats, img_feats, lengths)
        H = self.cfg.RNN_DIM
        h_t = torch.zeros(1, batch_size, H, device=self.device)
        c_t = torch.zeros(1, batch_size, H, device=self.device)
        ctx_mask = U.LengthMask(lengths, ctx.shape[1])
        if insts is None:
            insts = self.gt_words()
        insts = torch.as_tensor(insts, dtype=torch.int64, device=self.device)
        logits, _, _ = self.decoder(insts, ctx, ctx_mask, h_t, c_t)
        logits = logits.permute(0, 2, 1).contiguous()          # [B, V, L]
        if for_listener:
            return F.cross_entropy(logits[:, :, :-1], insts[:, 1:], ignore_index=self.pad, reduction="none")
        loss = F.cross_entropy(logits[:, :, :-1], insts[:, 1:], ignore_index=self.pad)
        if train:
            return loss
        _, predict = logits.max(dim=1)
        gt_mask = insts != self.pad
        correct = (predict[:, :-1] == insts[:, 1:]) & gt_mask[:, 1:]
        word_accu = correct.sum().item() / gt_mask[:, 1:].sum().item()
        sent_accu = (correct.sum(dim=1) == gt_mask[:, 1:].sum(dim=1)).sum().item() / batch_size
        return loss.item(), word_accu, sent_accu

    # ---- speaker.py:292-376 ------------------------------------------------------------------------------------------
    def infer_batch(self, sampling=False, train=False, featdropmask=None):
        """Greedy (or sampled) decoding of one instruction per path.  Not sampling: int64 [B, len] numpy; sampling and
        train: (words, log_probs [B, len], hidden states [B, len, H], entropies [B, len])."""
        self._mode(train)
        (img_feats, can_feats), lengths = self.from_shortest_path()
        B = can_feats.shape[0]
        if featdropmask is not None:                        # one mask per feature for the whole batch (envdrop back-translation)
            A = self.angle_feat_size
            can_feats = torch.cat((can_feats[..., :-A] * featdropmask, can_feats[..., -A:]), -1)
            img = ops.gather_pano(img_feats.store, img_feats.vp, img_feats.view)
            img_feats = torch.cat((img[..., :-A] * featdropmask, img[..., -A:]), -1).view(B, -1, ops.N_VIEWS, self.feature_size)
        ctx = self.encoder(can_feats, img_feats, lengths, already_dropfeat=(featdropmask is not None))
        ctx_mask = U.LengthMask(lengths, ctx.shape[1])
        H = self.cfg.RNN_DIM
        h_t = torch.zeros(1, B, H, device=self.device)
        c_t = torch.zeros(1, B, H, device=self.device)
        words, log_probs, hidden_states, entropies = [], [], [], []
        ended = torch.zeros(B, dtype=torch.bool, device=self.device)
        word = torch.full((B, 1), self.bos, dtype=torch.int64, device=self.device)
        for i in range(self.cfg.MAX_DECODE):
            logits, h_t, c_t = self.decoder(word, ctx, ctx_mask, h_t, c_t)
            logits = logits.reshape(B, -1).clone()
            logits[:, self.unk] = -float("inf")
            if sampling:
                m = torch.distributions.Categorical(F.softmax(logits, -1))
                w = m.sample()
                lp, ent, hs = m.log_prob(w), m.entropy(), h_t[0]
                if not train:
                    lp, ent, hs = lp.detach(), ent.detach(), hs.detach()
                log_probs.append(lp)
                hidden_states.append(hs)
                entropies.append(ent)
            else:
                w = logits.argmax(1)
            words.append(torch.where(ended, torch.full_like(w, self.pad), w))
            word = w.view(-1, 1)
            ended = ended | (words[-1] == self.eos)
            if (i % 8 == 7 or i == self.cfg.MAX_DECODE - 1) and bool(ended.all()):   # one read-back per 8 words
                break
        out = torch.stack(words, 1).cpu().numpy()
        # the reference stops at the first step after which every row has ended: trim what the sparser poll ran past it
        done = (out == self.eos).any(1)
        if done.all():
            last = int(np.max(np.argmax(out == self.eos, axis=1))) + 1
            out = out[:, :last]
            log_probs, hidden_states, entropies = log_probs[:last], hidden_states[:last], entropies[:last]
        if train and sampling:
            return out, torch.stack(log_probs, 1), torch.stack(hidden_states, 1), torch.stack(entropies, 1)
        return out

    # ---- speaker.py:378-413 ------------------------------------------------------------------------------------------
    def save(self, epoch, path):
        the_dir, _ = os.path.split(path)
        if the_dir:
            os.makedirs(the_dir, exist_ok=True)
        states = {name: {"epoch": epoch + 1, "state_dict": model.state_dict(), "optimizer": opt.state_dict()}
                  for name, model, opt in (("encoder", self.encoder, self.encoder_optimizer),
                                           ("decoder", self.decoder, self.decoder_optimizer))}
        torch.save(states, path)

    def load(self, path, cuda=0):
        states = torch.load(path, map_location=self.device)
        for name, model, opt in (("encoder", self.encoder, self.encoder_optimizer),
                                 ("decoder", self.decoder, self.decoder_optimizer)):
            state = model.state_dict()
            state.update(states[name]["state_dict"])
            model.load_state_dict(state)
            if self.cfg.LOAD_OPTIM:
                opt.load_state_dict(states[name]["optimizer"])
        return states["encoder"]["epoch"] - 1
