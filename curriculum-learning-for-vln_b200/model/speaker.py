"""Speaker (instruction generator) modules on the sm_100a kernels.

Same constructor arguments, ``forward`` signatures and ``state_dict`` keys as the reference's ``SpeakerEncoder``
(src/model/units.py:286-341) and ``SpeakerDecoder`` (:344-390), so checkpoints are interchangeable.  Where the arithmetic
runs:

  * the per-step panorama ``feature`` [B, T, 36, 2176] may be an ``ops.PanoView`` over the B*T (viewpoint, view) pairs of
    the path: gather + feature dropout + soft-dot attention are then the one fused kernel of the navigation agents
    (csrc/pano_attn.cu / pano_stream.cu), each table row read once — the reference materialises 36 x 2176 floats per step;
  * the (bi)directional LSTMs run on the persistent tcgen05 recurrence (csrc/lstm_tc.cu), their input projections and every
    nn.Linear on the tcgen05 bf16x3 GEMM; the decoder's attention over the path context is csrc/ctx_attn.cu;
  * dropout draws from the library's Philox stream.
"""
import torch
import torch.nn as nn

from .. import ops
from .units import KernelModule, SoftDotAttention, LengthMask


def _lstm_all_steps(module, lstm, x, sfx_list):
    """nn.LSTM(batch_first=True) over every step of x [B, T, D] from a zero state (no packing: the reference's speaker
    runs its LSTMs over the padded steps too, units.py:322, :366): -> (out [B, T, n_dir * H], h_last, c_last)."""
    B, T, _ = x.shape
    xprojs, whhs = [], []
    for sfx in sfx_list:
        w_ih, w_hh = getattr(lstm, "weight_ih" + sfx), getattr(lstm, "weight_hh" + sfx)
        bias = getattr(lstm, "bias_ih" + sfx) + getattr(lstm, "bias_hh" + sfx)
        xprojs.append(ops.linear(x.reshape(B * T, -1), w_ih, bias).view(B, T, -1))
        whhs.append(w_hh)
    lengths = torch.full((B,), T, dtype=torch.int32, device=x.device)
    H = whhs[0].shape[1]
    if H in ops.LSTM_KERNEL_H:
        return ops.lstm_layer(xprojs, whhs, lengths)
    res = [ops.lstm_sequence(xp, lengths, w, reverse=bool(d)) for d, (xp, w) in enumerate(zip(xprojs, whhs))]
    if len(res) == 1:
        return res[0]
    return torch.cat([r[0] for r in res], 2), torch.cat([r[1] for r in res], 1), torch.cat([r[2] for r in res], 1)


class SpeakerEncoder(KernelModule):
    def __init__(self, feature_size, hidden_size, dropout_ratio, bidirectional, angle_feat_size, feat_dropout):
        super().__init__()
        self.num_directions = 2 if bidirectional else 1
        self.hidden_size = hidden_size
        self.num_layers = 1
        self.feature_size = feature_size
        self.angle_feat_size = angle_feat_size
        self.dropout_ratio, self.feat_dropout = dropout_ratio, feat_dropout
        # parameter containers (names / shapes / init of the reference); forward never calls them
        self.lstm = nn.LSTM(feature_size, hidden_size // self.num_directions, self.num_layers, batch_first=True,
                            bidirectional=bidirectional)
        self.attention_layer = SoftDotAttention(query_dim=hidden_size, context_dim=feature_size)
        self.post_lstm = nn.LSTM(hidden_size, hidden_size // self.num_directions, self.num_layers, batch_first=True,
                                 bidirectional=bidirectional)

    def _sfx(self):
        return ["_l0", "_l0_reverse"][:self.num_directions]

    def forward(self, action_embeds, feature, lengths=None, already_dropfeat=False):
        """action_embeds [B, T, 2176]; feature: dense [B, T, 36, 2176] or an ops.PanoView over the B*T steps (row b*T + t).
        -> context [B, T, hidden]."""
        B, T, _ = action_embeds.shape
        A = self.angle_feat_size
        p_feat = self.feat_dropout if (self.training and not already_dropfeat) else 0.0
        x = action_embeds
        if p_feat > 0.0:                                   # not the spatial part (units.py:318-319)
            x = torch.cat((self._drop(x[..., :-A].contiguous(), p_feat, "spk_can"), x[..., -A:]), -1)
        ctx, _, _ = _lstm_all_steps(self, self.lstm, x, self._sfx())
        ctx = self._drop(ctx, self.dropout_ratio, "spk_ctx")
        q = ctx.reshape(B * T, self.hidden_size)
        if isinstance(feature, ops.PanoView):
            if p_feat > 0.0:                               # feature dropout inside the gather kernel (units.py:328-329)
                rng = self._rng(q)
                feature.drop_p, feature.rng = p_feat, rng
                feature.call_off = rng.next("spk_img", (B * T, ops.N_VIEWS, ops.F_DIM - A), p_feat)
        else:
            feature = feature.reshape(B * T, -1, self.feature_size)
            if p_feat > 0.0:
                feature = torch.cat((self._drop(feature[..., :-A].contiguous(), p_feat, "spk_img"), feature[..., -A:]), -1)
        x, _ = self.attention_layer(q, feature)
        x = self._drop(x.view(B, T, -1), self.dropout_ratio, "spk_att")
        x, _, _ = _lstm_all_steps(self, self.post_lstm, x, self._sfx())
        return self._drop(x, self.dropout_ratio, "spk_post")


class SpeakerDecoder(KernelModule):
    def __init__(self, vocab_size, embedding_size, padding_idx, hidden_size, dropout_ratio):
        super().__init__()
        self.hidden_size = hidden_size
        self.dropout_ratio = dropout_ratio
        self.embedding = nn.Embedding(vocab_size, embedding_size, padding_idx)
        self.lstm = nn.LSTM(embedding_size, hidden_size, batch_first=True)
        self.attention_layer = SoftDotAttention(query_dim=hidden_size)
        self.projection = nn.Linear(hidden_size, vocab_size)
        self.baseline_projection = nn.Sequential(nn.Linear(hidden_size, 128), nn.ReLU(), nn.Dropout(dropout_ratio),
                                                 nn.Linear(128, 1))

    def forward(self, words, ctx, ctx_mask, h0, c0):
        """words int64 [Bw, Lw]; ctx [B, T, H] with Bw a multiple of B (beam search repeats contexts); ctx_mask: boolean
        [B, T] (True = padded step) or a LengthMask; h0 / c0 [1, Bw, H].  -> (logit [Bw, Lw, V], h1, c1)."""
        Bw, Lw = words.shape
        H = self.hidden_size
        p = self.dropout_ratio if self.training else 0.0
        embeds = ops.embed_dropout(words, self.embedding.weight, self.embedding.padding_idx, p, self._rng(words), "spk_emb")
        w_ih, w_hh = self.lstm.weight_ih_l0, self.lstm.weight_hh_l0
        b_ih, b_hh = self.lstm.bias_ih_l0, self.lstm.bias_hh_l0
        zero_state = not (bool(h0.any()) or bool(c0.any())) if Lw > 1 else False
        if Lw > 1 and zero_state:
            x, h1, c1 = _lstm_all_steps(self, self.lstm, embeds, ["_l0"])
        else:                                              # step-wise decoding carries the state (speaker.py:337-340)
            h, c = h0[0], c0[0]
            outs = []
            for t in range(Lw):
                h, c = ops.lstm_cell(embeds[:, t], h, c, w_ih, w_hh, b_ih, b_hh)
                outs.append(h)
            x, h1, c1 = torch.stack(outs, 1), h, c
        x = self._drop(x, self.dropout_ratio, "spk_dec")
        B, T, _ = ctx.shape
        mult = (Bw * Lw) // B                              # words per context (units.py:372-373)
        lens = ctx_mask.lengths if isinstance(ctx_mask, LengthMask) else ops.mask_to_lengths(ctx_mask, T)
        ctx_rep = ctx.unsqueeze(1).expand(-1, mult, -1, -1).reshape(Bw * Lw, T, H)
        lens_rep = lens.unsqueeze(1).expand(-1, mult).reshape(Bw * Lw)
        x, _ = self.attention_layer(x.reshape(Bw * Lw, H), ctx_rep, LengthMask(lens_rep, T))
        x = self._drop(x.view(Bw, Lw, H), self.dropout_ratio, "spk_out")
        logit = ops.linear(x.reshape(Bw * Lw, H), self.projection.weight, self.projection.bias).view(Bw, Lw, -1)
        return logit, h1.unsqueeze(0), c1.unsqueeze(0)
