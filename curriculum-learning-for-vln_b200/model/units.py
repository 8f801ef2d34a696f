"""Instruction encoder and attention / scoring blocks on the sm_100a kernels.

Same constructor arguments, ``forward`` signatures and ``state_dict`` keys as the reference's
``src/model/units.py`` (EncoderLSTM :12-74, SoftDotAttention :77-122, VisualSoftDotAttention
:125-160, ActionScoring :163-185, PositionalEncoding :188-207, MLPwithBN :210-242), so a
checkpoint of either side loads into the other.  What differs is where the arithmetic runs:

  * the feature arguments may be lazy views of the HBM table (ops.PanoView / ops.CandView)
    instead of materialised [B,36,2176] / [B,C,2176] tensors — then gather, feature dropout,
    dot products, softmax and the weighted sum are ONE kernel reading each table row once;
  * linear layers that only feed a dot product with table rows are folded algebraically
    (x.(W t) instead of (W x).t), which removes the (B*36)x2176x256 key projection of the
    Follower's visual attention and the per-candidate projection of ActionScoring;
  * packed-sequence LSTMs run as length-masked recurrences;
  * dropout draws from the library's Philox stream (ops.Rng), not from torch's generator.
"""
import math

import torch
import torch.nn as nn

from .. import ops

_global_rng = {}


def default_rng(device):
    key = str(device)
    if key not in _global_rng:
        _global_rng[key] = ops.Rng(2020, device)
    return _global_rng[key]


class KernelModule(nn.Module):
    """nn.Module whose dropout sites draw from a shared ops.Rng (set with ``use_rng``)."""
    rng = None

    def _rng(self, ref):
        return self.rng if self.rng is not None else default_rng(ref.device)

    def _drop(self, x, p, tag):
        if not self.training or p <= 0.0:
            return x
        return ops.dropout(x, p, self._rng(x), tag)


def use_rng(module, rng):
    for m in module.modules():
        if isinstance(m, KernelModule):
            m.rng = rng
    return module


class LengthMask:
    """A suffix pad mask given by its lengths (what `seq == pad` / length2mask produce,
    base.py:128, misc.py:481-486).  ``dense()`` gives the boolean tensor the reference passes."""

    def __init__(self, lengths, L):
        self.lengths = lengths if lengths.dtype == torch.int32 else lengths.to(torch.int32)
        self.L = L

    def dense(self):
        return torch.arange(self.L, device=self.lengths.device).unsqueeze(0) >= self.lengths.unsqueeze(1)


def _lengths_of(mask, B, L, device):
    if mask is None:
        return torch.full((B,), L, dtype=torch.int32, device=device)
    if isinstance(mask, LengthMask):
        return mask.lengths
    return ops.mask_to_lengths(mask, L)


class PhiloxDropout(KernelModule):
    """nn.Dropout stand-in for nn.Sequential containers (keeps the child indices, hence the
    state_dict keys, of MLPwithBN)."""

    def __init__(self, p, tag="mlp"):
        super().__init__()
        self.p, self.tag = p, tag

    def forward(self, x):
        return self._drop(x, self.p, self.tag)


class KernelLinear(nn.Linear):
    """nn.Linear whose forward runs on the tcgen05 bf16x3 kernel (ops.linear: skinny or tall launch by row count);
    same parameters / state_dict keys, so it can sit inside the nn.Sequential of MLPwithBN (units.py:218-236 — the
    (B*C) x 2176 x 1024 projection that dominates a Self-Monitor step)."""

    def forward(self, x):
        y = ops.linear(x.reshape(-1, x.shape[-1]), self.weight, self.bias)
        return y.view(*x.shape[:-1], self.weight.shape[0])


class EncoderLSTM(KernelModule):
    def __init__(self, vocab_size, embed_size, hidden_size, padding_idx, drop_ratio=0.5, bidirectional=False,
                 num_layers=1, glove=None):
        super().__init__()
        self.embed_size, self.vocab_size = embed_size, vocab_size
        self.num_directions = 2 if bidirectional else 1
        self.hidden_size = hidden_size // self.num_directions
        self.num_layers = num_layers
        self.drop_ratio = drop_ratio
        self.use_glove = glove is not None
        if self.use_glove:
            self.embedding = nn.Embedding.from_pretrained(torch.from_numpy(glove), freeze=True)
        else:
            self.embedding = nn.Embedding(vocab_size, embed_size, padding_idx=padding_idx)
        # parameter containers only (names/shapes/init of the reference); forward never calls them
        self.lstm = nn.LSTM(embed_size, self.hidden_size, dropout=drop_ratio * (num_layers > 1),
                            num_layers=num_layers, batch_first=True, bidirectional=bidirectional)
        self.enc2dec = nn.Linear(self.hidden_size * self.num_directions, self.hidden_size * self.num_directions)

    def forward(self, inputs, lengths, already_sorted=True):
        """inputs int64 [B,L], lengths [B] (device tensor preferred; a host tensor is copied).
        Returns (ctx [B,L,H], decoder_init [B,H], c_t [B,H])."""
        dev = inputs.device
        if lengths.device != dev:
            lengths = lengths.to(dev)
        if self.use_glove:
            x = self.embedding(inputs)                       # frozen GloVe rows, no dropout (units.py:49-52)
        else:
            p = self.drop_ratio if self.training else 0.0
            x = ops.embed_dropout(inputs, self.embedding.weight, self.embedding.padding_idx, p, self._rng(inputs))
        B, L, _ = x.shape
        h_last = c_last = None
        for layer in range(self.num_layers):
            xprojs, whhs = [], []
            for d in range(self.num_directions):
                sfx = f"_l{layer}" + ("_reverse" if d else "")
                w_ih, w_hh = getattr(self.lstm, "weight_ih" + sfx), getattr(self.lstm, "weight_hh" + sfx)
                bias = getattr(self.lstm, "bias_ih" + sfx) + getattr(self.lstm, "bias_hh" + sfx)
                # layer 0's dx only feeds the embedding gradient: TF32 library GEMM (see ops._LinearTall)
                xprojs.append(ops.linear(x.reshape(B * L, -1), w_ih, bias, dx_tf32=(layer == 0)).view(B, L, -1))
                whhs.append(w_hh)
            if self.hidden_size in ops.LSTM_KERNEL_H:
                # persistent cluster kernel (tcgen05, W_hh resident in tensor memory), both directions in one launch:
                # 128 / 256 per direction (Follower, EnvDrop) and 512 (Self-Monitor's uni-directional encoder)
                x, h_last, c_last = ops.lstm_layer(xprojs, whhs, lengths)
            else:
                # other hidden sizes: step-by-step recurrence (one pointwise kernel + one GEMM per timestep)
                res = [ops.lstm_sequence(xp, lengths, w, reverse=bool(d)) for d, (xp, w) in enumerate(zip(xprojs, whhs))]
                x = torch.cat([r[0] for r in res], 2) if len(res) > 1 else res[0][0]
                h_last = torch.cat([r[1] for r in res], 1) if len(res) > 1 else res[0][1]
                c_last = torch.cat([r[2] for r in res], 1) if len(res) > 1 else res[0][2]
            if layer + 1 < self.num_layers:
                x = self._drop(x, self.drop_ratio, "enc_interlayer")
        decoder_init = torch.tanh(ops.linear(h_last, self.enc2dec.weight, self.enc2dec.bias))
        ctx = self._drop(x, self.drop_ratio, "enc_ctx")
        return ctx, decoder_init, c_last


def _attend(target, context, mask):
    """softmax(context . target) and the weighted context, for a table view or a dense tensor."""
    if isinstance(context, ops.PanoView):
        return ops.pano_attn(context.store, context.vp, context.view, target, getattr(context, "drop_p", 0.0),
                             getattr(context, "rng", None), getattr(context, "call_off", 0),
                             getattr(context, "split", 1))
    B, S, _ = context.shape
    return ops.ctx_attn(context, target, _lengths_of(mask, B, S, context.device))


class SoftDotAttention(KernelModule):
    def __init__(self, query_dim, context_only=False, context_dim=None):
        super().__init__()
        self.context_only = context_only
        ctx_dim = query_dim if context_dim is None else context_dim
        self.linear_in = nn.Linear(query_dim, ctx_dim, bias=False)
        if not context_only:
            self.linear_out = nn.Linear(query_dim + ctx_dim, query_dim, bias=False)

    def forward(self, h, context, mask=None):
        target = ops.linear(h, self.linear_in.weight)
        weighted, attn = _attend(target, context, mask)
        if self.context_only:
            return weighted, attn
        h_tilde = torch.tanh(ops.linear(torch.cat((weighted, h), 1), self.linear_out.weight))
        return h_tilde, attn


class VisualSoftDotAttention(KernelModule):
    def __init__(self, h_dim, v_dim=None, dot_dim=256):
        super().__init__()
        self.linear_in_h = nn.Linear(h_dim, dot_dim, bias=True)
        self.use_v_linear = v_dim is not None
        if self.use_v_linear:
            self.linear_in_v = nn.Linear(v_dim, dot_dim, bias=True)

    def forward(self, h, visual_context, mask=None):
        target = ops.linear(h, self.linear_in_h.weight, self.linear_in_h.bias)
        if self.use_v_linear:
            # (W_v x_s + b_v) . t  =  x_s . (W_v^T t) + b_v . t ; the second term is constant over s and
            # cancels in the softmax, so the key projection never has to be computed per view.
            target = target @ self.linear_in_v.weight
        return _attend(target, visual_context, mask)


class ActionScoring(KernelModule):
    def __init__(self, action_size, hidden_size, dot_size=256):
        super().__init__()
        self.linear_act = nn.Linear(action_size, dot_size, bias=True)
        self.linear_hid = nn.Linear(hidden_size, dot_size, bias=True)
        self.linear_out = nn.Linear(dot_size, 1, bias=True)

    def forward(self, act_cands, h_tilde):
        # logit_j = w_out . ((W_a x_j + b_a) * t) + b_out = x_j . (W_a^T (w_out * t)) + b_a . (w_out * t) + b_out
        t = ops.linear(h_tilde, self.linear_hid.weight, self.linear_hid.bias) * self.linear_out.weight
        tgt = t @ self.linear_act.weight
        bias = t @ self.linear_act.bias + self.linear_out.bias
        if isinstance(act_cands, ops.CandView):
            return ops.cand_logits(act_cands.store, act_cands.vp, act_cands.view, tgt, bias)
        return torch.bmm(act_cands, tgt.unsqueeze(2)).squeeze(2) + bias.unsqueeze(1)


class PositionalEncoding(KernelModule):
    def __init__(self, d_model, dropout, max_len=80):
        super().__init__()
        self.p = dropout
        pe = torch.zeros(max_len, d_model)
        position = torch.arange(0, max_len).float().unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2).float() * -(math.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe.unsqueeze(0))

    def forward(self, x):
        return self._drop(x + self.pe[:, :x.size(1)], self.p, "pos")


class MLPwithBN(KernelModule):
    def __init__(self, input_size, hidden_size, out_size=None, dropout=0.0, use_bn=False, use_bias=True, relu=True):
        super().__init__()
        self.in_size = input_size
        layers = []
        if use_bn:
            layers.append(nn.BatchNorm1d(input_size))
        dims = [input_size] + list(hidden_size)
        for d_in, d_out in zip(dims[:-1], dims[1:]):
            layers.append(KernelLinear(d_in, d_out, bias=use_bias))
            if use_bn:
                layers.append(nn.BatchNorm1d(d_out))
            if dropout > 0:
                layers.append(PhiloxDropout(dropout))
            if relu:
                layers.append(nn.ReLU(inplace=True))
        self.out_size = hidden_size[-1]
        if out_size:
            layers.append(KernelLinear(dims[-1], out_size, bias=use_bias))
            self.out_size = out_size
        self.mlp = nn.Sequential(*layers)

    def forward(self, x, row_weight=None, n_rows=None):
        """``row_weight`` [rows] (0/1) + ``n_rows`` (0-dim tensor = its sum): the BatchNorm statistics run over the
        weighted rows only.  The reference pads candidates to the batch's widest row and normalises over exactly those
        B x C rows (policy.py:144-149); here the candidate tensor always has 16 slots and the width C is a DEVICE value,
        so no host read-back sits in the rollout (the statistics, running averages included, are the same numbers)."""
        if row_weight is None:
            return self.mlp(x)
        for layer in self.mlp:
            if isinstance(layer, nn.BatchNorm1d) and layer.training:
                x = _masked_batch_norm(layer, x, row_weight, n_rows)
            else:
                x = layer(x)
        return x


def _masked_batch_norm(bn, x, w, n):
    """nn.BatchNorm1d in training mode over the rows with weight 1: batch mean / biased variance for the output,
    running_mean / running_var (unbiased, momentum) / num_batches_tracked updated as torch does."""
    wc = w.unsqueeze(1)
    mean = (x * wc).sum(0) / n
    d = (x - mean) * wc
    var = (d * d).sum(0) / n
    with torch.no_grad():
        m = bn.momentum if bn.momentum is not None else 0.1
        # (.data: like the fused batch-norm kernels, the update must not bump the buffers' autograd version counters —
        #  the standard BatchNorm call on the previous-action rows has saved them for its backward)
        bn.running_mean.data.mul_(1 - m).add_(mean.detach() * m)
        bn.running_var.data.mul_(1 - m).add_(var.detach() * (n / (n - 1).clamp(min=1)) * m)
        bn.num_batches_tracked.data.add_(1)
    return (x - mean) * torch.rsqrt(var + bn.eps) * bn.weight + bn.bias
