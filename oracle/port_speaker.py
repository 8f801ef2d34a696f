"""Oracle port of the Speaker (instruction generator) — TEST INFRASTRUCTURE ONLY.

CPU / fp32 restatement of the reference's SpeakerEncoder (src/model/units.py:286-341), SpeakerDecoder (:344-390) and of
the Speaker's path features, teacher forcing and greedy / sampled decoding (src/agent/speaker.py:191-226, :235-290,
:292-376), on the reference's state_dicts and on any environment with the obs-dict protocol of oracle/port_env.py
(``reset() / observe() / step(actions, obs)``; keys "candidates", "viewpointId", "teacher", "feature", ...).

Pinned by tests/_ref_check_speaker.py against the UNMODIFIED reference run in this container: the real modules
(eval, and train with torch-RNG-matched dropout), the real ``Speaker.teacher_forcing(features=..., insts=...)`` entry
point beam search uses, and the real ``from_shortest_path`` / ``infer_batch`` driven through an adapter that gives the
reference's own R2RBatch the EnvDrop-original observation keys the Speaker class expects (the class as shipped reads
``ob['candidate']`` / ``ob['viewpoint']``, which the repository's env never produces — speaker.py:165-189 vs
common_env.py:299-330).

Dropout: ``drop`` is a port_modules.Drop (None = eval); tags "spk_can", "spk_ctx", "spk_img", "spk_att", "spk_post",
"spk_emb", "spk_dec", "spk_out" name the sites in call order, so a CUDA run's Philox masks can be injected.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import port_modules as P
from .port_env import angle_feat


# ---- nn.LSTM(batch_first=True) over every step from a zero state (no packing) ----------------------------------------
def lstm_all_steps(sd, pfx, x, bidirectional, h0=None, c0=None):
    """-> (out [B, T, n_dir * H], h_n [n_dir, B, H], c_n).  ``h0`` / ``c0`` [n_dir, B, H] (default zero)."""
    B, T, _ = x.shape
    outs, hs, cs = [], [], []
    for d, sfx in enumerate(["_l0", "_l0_reverse"][:2 if bidirectional else 1]):
        w_ih, w_hh = sd[f"{pfx}.weight_ih{sfx}"], sd[f"{pfx}.weight_hh{sfx}"]
        b_ih, b_hh = sd[f"{pfx}.bias_ih{sfx}"], sd[f"{pfx}.bias_hh{sfx}"]
        H = w_hh.shape[1]
        h = x.new_zeros(B, H) if h0 is None else h0[d]
        c = x.new_zeros(B, H) if c0 is None else c0[d]
        row = [None] * T
        for t in (range(T - 1, -1, -1) if d else range(T)):
            h, c = P.lstm_cell(x[:, t], h, c, w_ih, w_hh, b_ih, b_hh)
            row[t] = h
        outs.append(torch.stack(row, 0))
        hs.append(h)
        cs.append(c)
    # memory layout of nn.LSTM(batch_first=True)'s output: [T, B, .] contiguous, viewed transposed — torch's dropout draws
    # its mask in memory order, so the layout is part of what "the same torch RNG stream" means for the pin against the
    # real modules (tests/_ref_check_speaker.py)
    return torch.cat(outs, 2).transpose(0, 1), torch.stack(hs, 0), torch.stack(cs, 0)


def speaker_encoder(sd, action_embeds, feature, *, hidden=512, bidirectional=True, angle=128, p=0.6, pf=0.3, drop=None,
                    already_dropfeat=False):
    """SpeakerEncoder.forward (units.py:311-341): action_embeds [B, T, F], feature [B, T, 36, F] -> ctx [B, T, hidden]."""
    B, T, Fdim = action_embeds.shape
    x = action_embeds
    use_fd = drop is not None and not already_dropfeat
    if use_fd:
        x = torch.cat((P._drop(drop, x[..., :-angle], pf, "spk_can"), x[..., -angle:]), -1)
    ctx, _, _ = lstm_all_steps(sd, "lstm", x, bidirectional)
    ctx = P._drop(drop, ctx, p, "spk_ctx")
    feat = feature.reshape(B * T, -1, Fdim)
    if use_fd:
        feat = torch.cat((P._drop(drop, feat[..., :-angle], pf, "spk_img"), feat[..., -angle:]), -1)
    x, _ = P.soft_dot_attention(ctx.reshape(B * T, hidden), feat, sd["attention_layer.linear_in.weight"],
                                sd["attention_layer.linear_out.weight"])
    x = P._drop(drop, x.view(B, T, -1), p, "spk_att")
    x, _, _ = lstm_all_steps(sd, "post_lstm", x, bidirectional)
    return P._drop(drop, x, p, "spk_post")


def speaker_decoder(sd, words, ctx, ctx_mask, h0, c0, *, hidden=512, p=0.6, drop=None, pad=0):
    """SpeakerDecoder.forward (units.py:364-390): words [Bw, Lw] -> (logit [Bw, Lw, V], h1 [1, Bw, H], c1)."""
    Bw, Lw = words.shape
    embeds = F.embedding(words, sd["embedding.weight"], padding_idx=pad)   # (no gradient into the <PAD> row)
    embeds = P._drop(drop, embeds, p, "spk_emb")
    x, h1, c1 = lstm_all_steps(sd, "lstm", embeds, False, h0, c0)
    x = P._drop(drop, x, p, "spk_dec")
    n = Bw * Lw
    mult = n // ctx.shape[0]
    x, _ = P.soft_dot_attention(
        x.reshape(n, hidden), ctx.unsqueeze(1).expand(-1, mult, -1, -1).reshape(n, -1, hidden),
        sd["attention_layer.linear_in.weight"], sd["attention_layer.linear_out.weight"],
        mask=ctx_mask.unsqueeze(1).expand(-1, mult, -1).reshape(n, -1))
    x = P._drop(drop, x.view(Bw, Lw, hidden), p, "spk_out")
    return F.linear(x, sd["projection.weight"], sd["projection.bias"]), h1, c1


# ---- the Speaker's use of the environment (speaker.py:160-226) -----------------------------------------------------
def from_shortest_path(env, obs, fdim=2176, angle=128, get_first_feat=False):
    """Follows the teacher from the current observations ``obs``.  -> ((img_feats [B, T, 36, F], can_feats [B, T, F]
    [, first_feat]), lengths int64 [B], viewpoints per episode)."""
    B = len(obs)
    ended = np.zeros(B, bool)
    length = np.zeros(B, np.int64)
    img_feats, can_feats, viewpoints = [], [], [[] for _ in range(B)]
    first = np.zeros((B, fdim), np.float32)
    for i, ob in enumerate(obs):
        first[i, -angle:] = angle_feat(ob["heading"], ob["elevation"], angle)
    while not ended.all():
        for i, ob in enumerate(obs):
            viewpoints[i].append(ob["viewpointId"])
        img_feats.append(torch.from_numpy(np.stack([ob["feature"] for ob in obs]).astype(np.float32)))
        act = np.full(B, -1, np.int64)
        can = np.zeros((B, fdim), np.float32)
        for i, ob in enumerate(obs):
            if ended[i]:
                continue
            for k, c in enumerate(ob["candidates"]):
                if c["nextViewpointId"] == ob["teacher"]:
                    act[i] = k
                    can[i] = c["feature"]
                    break
            else:
                assert ob["teacher"] == ob["viewpointId"]              # "stay here": the STOP action, zero feature
        can_feats.append(torch.from_numpy(can))
        obs = env.step(act, obs)
        length += (1 - ended)
        ended |= act == -1
    feats = (torch.stack(img_feats, 1).contiguous(), torch.stack(can_feats, 1).contiguous())
    if get_first_feat:
        feats = feats + (torch.from_numpy(first),)
    return feats, length, viewpoints


def length_mask(lengths, size=None):                                     # misc.py:481-486
    lengths = torch.as_tensor(np.asarray(lengths), dtype=torch.int64)
    size = int(lengths.max()) if size is None else size
    return torch.arange(size).unsqueeze(0) > (lengths - 1).unsqueeze(1)


class SpeakerPort:
    def __init__(self, encoder_sd, decoder_sd, *, hidden=512, bidirectional=True, p=0.6, pf=0.3, angle=128, pad=0, unk=1,
                 eos=2, bos=3, max_decode=120):
        self.enc, self.dec = encoder_sd, decoder_sd
        self.hidden, self.bi, self.p, self.pf, self.angle = hidden, bidirectional, p, pf, angle
        self.pad, self.unk, self.eos, self.bos, self.max_decode = pad, unk, eos, bos, max_decode

    def encode(self, features, drop=None, already_dropfeat=False):
        (img_feats, can_feats), lengths = features
        ctx = speaker_encoder(self.enc, can_feats, img_feats, hidden=self.hidden, bidirectional=self.bi, angle=self.angle,
                              p=self.p, pf=self.pf, drop=drop, already_dropfeat=already_dropfeat)
        return ctx, length_mask(lengths, ctx.shape[1])

    def teacher_forcing(self, features, insts, *, train=True, for_listener=False, drop=None):
        """speaker.py:235-290 with the features given.  train: loss; for_listener: per-word CE [B, L-1];
        not train: (loss, word accuracy, sentence accuracy, logits)."""
        ctx, ctx_mask = self.encode(features, drop)
        B = ctx.shape[0]
        z = torch.zeros(1, B, self.hidden)
        logits, _, _ = speaker_decoder(self.dec, insts, ctx, ctx_mask, z, z, hidden=self.hidden, p=self.p, drop=drop)
        lg = logits.permute(0, 2, 1).contiguous()
        if for_listener:
            return F.cross_entropy(lg[:, :, :-1], insts[:, 1:], ignore_index=self.pad, reduction="none")
        loss = F.cross_entropy(lg[:, :, :-1], insts[:, 1:], ignore_index=self.pad)
        if train:
            return loss
        _, predict = lg.max(dim=1)
        gt_mask = insts != self.pad
        correct = (predict[:, :-1] == insts[:, 1:]) & gt_mask[:, 1:]
        word_accu = correct.sum().item() / gt_mask[:, 1:].sum().item()
        sent_accu = (correct.sum(dim=1) == gt_mask[:, 1:].sum(dim=1)).sum().item() / B
        return loss.item(), word_accu, sent_accu, logits

    def infer_batch(self, features, forced=None):
        """Greedy decoding (speaker.py:292-376, sampling=False); ``forced`` [B, len] replays another run's words as the
        inputs of the following steps.  -> (words int64 [B, len], logits per step [len, B, V])."""
        ctx, ctx_mask = self.encode(features)
        B = ctx.shape[0]
        h = torch.zeros(1, B, self.hidden)
        c = torch.zeros(1, B, self.hidden)
        ended = np.zeros(B, bool)
        word = torch.full((B, 1), self.bos, dtype=torch.int64)
        words, steps = [], []
        for i in range(self.max_decode):
            logits, h, c = speaker_decoder(self.dec, word, ctx, ctx_mask, h, c, hidden=self.hidden, p=self.p)
            logits = logits.reshape(B, -1).clone()
            logits[:, self.unk] = -float("inf")
            steps.append(logits)
            w = logits.argmax(1)
            if forced is not None:
                if i >= forced.shape[1]:
                    break
                w = torch.as_tensor(forced[:, i], dtype=torch.int64)
            cpu = w.numpy().copy()
            cpu[ended] = self.pad
            words.append(cpu)
            word = w.view(-1, 1)
            ended = ended | (cpu == self.eos)
            if forced is None and ended.all():
                break
        return np.stack(words, 1), torch.stack(steps[:len(words)], 0)
