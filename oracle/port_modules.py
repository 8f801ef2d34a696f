"""Oracle port of the reference's model math — TEST INFRASTRUCTURE ONLY.

Functional fp32 PyTorch restatement (weights passed as the reference's own
``state_dict`` mappings, so a checkpoint of either side feeds both).  Every
function cites the reference lines it follows (paths relative to
/root/reference/tasks/R2R-judy/).  Pinned against the reference by
tests/test_oracle_vs_reference.py (container) and tests/golden/*.pt (anywhere).

Dropout: the reference uses nn.Dropout everywhere.  ``Drop`` reproduces it in
three ways: ``None`` = eval (identity); ``Drop("torch")`` = F.dropout on the
global RNG, in the reference's call order (bit-identical masks on CPU);
``Drop(masks={tag: [mask, ...]})`` = injected keep-masks, consumed in order per
tag, scaled by 1/(1-p) — how the CUDA path's Philox masks are fed to the oracle.
"""
import math

import torch
import torch.nn.functional as F


class Drop:
    def __init__(self, mode="torch", masks=None):
        self.mode = "masks" if masks is not None else mode
        self.masks = {k: list(v) for k, v in (masks or {}).items()}

    def __call__(self, x, p, tag):
        if p <= 0.0:
            return x
        if self.mode == "torch":
            return F.dropout(x, p, True)
        keep = self.masks[tag].pop(0).to(x.dtype)
        return x * keep * (1.0 / (1.0 - p))


def _drop(drop, x, p, tag):
    return x if drop is None else drop(x, p, tag)


# --------------------------------------------------------------------------
# LSTM pieces (torch.nn.LSTM / LSTMCell semantics: gate order i, f, g, o)
# --------------------------------------------------------------------------
def lstm_cell(x, h, c, w_ih, w_hh, b_ih, b_hh):
    """nn.LSTMCell as used at policy.py:53, :159, :238."""
    gates = F.linear(x, w_ih, b_ih) + F.linear(h, w_hh, b_hh)
    i, f, g, o = gates.chunk(4, dim=1)
    c1 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    h1 = torch.sigmoid(o) * torch.tanh(c1)
    return h1, c1


def _lstm_direction(x, lengths, w_ih, w_hh, b_ih, b_hh, reverse):
    """One direction of a packed-sequence LSTM restated with length masks:
    rows stop updating at their own length (reverse rows start there), outputs
    past the length are zero.  Equivalent to pack_padded_sequence -> nn.LSTM ->
    pad_packed_sequence at units.py:58-71."""
    B, L, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    xin = F.linear(x, w_ih, b_ih)                       # hoisted input projection
    steps = range(L - 1, -1, -1) if reverse else range(L)
    lengths = lengths.to(x.device)
    outs = [None] * L
    for t in steps:
        live = (lengths > t).unsqueeze(1)
        gates = xin[:, t] + F.linear(h, w_hh, b_hh)
        i, f, g, o = gates.chunk(4, dim=1)
        c_new = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h_new = torch.sigmoid(o) * torch.tanh(c_new)
        h = torch.where(live, h_new, h)
        c = torch.where(live, c_new, c)
        outs[t] = torch.where(live, h_new, torch.zeros_like(h_new))
    return torch.stack(outs, dim=0), h, c               # time-major [L, B, H]


def encoder_lstm(sd, tokens, lengths, *, bidirectional, num_layers, drop_ratio,
                 drop=None):
    """EncoderLSTM.forward, units.py:48-74.  Returns (ctx, decoder_init, c_t)."""
    x = F.embedding(tokens, sd["embedding.weight"])     # padding row is zero by init
    x = _drop(drop, x, drop_ratio, "enc_embed")
    dirs = ("", "_reverse") if bidirectional else ("",)
    h_last = c_last = None
    for layer in range(num_layers):
        outs, hs, cs = [], [], []
        for d in dirs:
            sfx = f"_l{layer}{d}"
            o, h, c = _lstm_direction(
                x, lengths, sd["lstm.weight_ih" + sfx], sd["lstm.weight_hh" + sfx],
                sd["lstm.bias_ih" + sfx], sd["lstm.bias_hh" + sfx], reverse=bool(d))
            outs.append(o), hs.append(h), cs.append(c)
        # time-major storage viewed batch-first: the memory layout pad_packed_sequence(batch_first=True)
        # returns, so Drop("torch") draws the same mask as the reference's nn.Dropout on ctx
        x = torch.cat(outs, dim=2).transpose(0, 1)
        h_last, c_last = torch.cat(hs, dim=1), torch.cat(cs, dim=1)
        if layer + 1 < num_layers and drop is not None:  # nn.LSTM inter-layer dropout
            if drop.mode == "torch":                    # nn.LSTM draws it on the packed [sum(len), H] data
                from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence
                pk = pack_padded_sequence(x, lengths.cpu(), batch_first=True)
                pk = pk._replace(data=drop(pk.data, drop_ratio, "enc_interlayer"))
                x, _ = pad_packed_sequence(pk, batch_first=True, total_length=x.shape[1])
            else:
                x = drop(x, drop_ratio, "enc_interlayer")
    dec_init = torch.tanh(F.linear(h_last, sd["enc2dec.weight"], sd["enc2dec.bias"]))
    ctx = _drop(drop, x, drop_ratio, "enc_ctx")
    return ctx, dec_init, c_last


# --------------------------------------------------------------------------
# attention / scoring blocks
# --------------------------------------------------------------------------
def soft_dot_attention(h, context, w_in, w_out=None, mask=None):
    """SoftDotAttention.forward, units.py:100-122 (context_only iff w_out is None)."""
    target = F.linear(h, w_in)
    logit = torch.einsum("bsd,bd->bs", context, target)
    if mask is not None:
        logit = logit.masked_fill(mask, -math.inf)
    attn = torch.softmax(logit, dim=1)
    weighted = torch.einsum("bs,bsd->bd", attn, context)
    if w_out is None:
        return weighted, attn
    return torch.tanh(F.linear(torch.cat((weighted, h), 1), w_out)), attn


def visual_soft_dot_attention(h, visual, w_h, b_h, w_v=None, b_v=None, mask=None):
    """VisualSoftDotAttention.forward, units.py:138-160."""
    target = F.linear(h, w_h, b_h)
    keys = visual if w_v is None else F.linear(visual, w_v, b_v)
    logit = torch.einsum("bsd,bd->bs", keys, target)
    if mask is not None:
        logit = logit.masked_fill(mask, -math.inf)
    attn = torch.softmax(logit, dim=1)
    return torch.einsum("bs,bsd->bd", attn, visual), attn


def action_scoring(sd, pfx, cands, h_tilde):
    """ActionScoring.forward, units.py:173-185."""
    t = F.linear(h_tilde, sd[pfx + "linear_hid.weight"], sd[pfx + "linear_hid.bias"])
    k = F.linear(cands, sd[pfx + "linear_act.weight"], sd[pfx + "linear_act.bias"])
    return F.linear(k * t.unsqueeze(1), sd[pfx + "linear_out.weight"],
                    sd[pfx + "linear_out.bias"]).squeeze(2)


def mlp_with_bn(sd, pfx, x, training, drop=None, tag="mlp"):
    """MLPwithBN (BN -> Linear -> BN -> Dropout(0.5) -> ReLU), units.py:210-242,
    as built at policy.py:84-86.  Running stats in ``sd`` are updated in place."""
    def bn(x, i):
        k = f"{pfx}mlp.{i}."
        if training and (k + "num_batches_tracked") in sd:
            sd[k + "num_batches_tracked"] += 1
        return F.batch_norm(x, sd[k + "running_mean"], sd[k + "running_var"],
                            sd[k + "weight"], sd[k + "bias"], training, 0.1, 1e-5)
    x = bn(x, 0)
    x = F.linear(x, sd[pfx + "mlp.1.weight"], sd[pfx + "mlp.1.bias"])
    x = bn(x, 2)
    if training:
        x = _drop(drop, x, 0.5, tag)
    return torch.relu(x)


# --------------------------------------------------------------------------
# decoders (one step each)
# --------------------------------------------------------------------------
def envdrop_decoder(sd, a_prev, img, cand, h_tilde_prev, c0, ctx, ctx_mask, *,
                    drop_ratio=0.5, feat_drop_ratio=0.3, drop=None, angle=128):
    """EnvDropDecoder.forward, policy.py:208-246 (already_dropfeat=False).
    Unlike the reference it does NOT mutate img/cand; the dropped tensors are
    what the attention and the candidate logits see, as there."""
    act = torch.tanh(F.linear(a_prev, sd["act_embed.0.weight"], sd["act_embed.0.bias"]))
    act = _drop(drop, act, drop_ratio, "act")
    if drop is not None:
        nimg = img.shape[-1] - angle
        img = torch.cat((_drop(drop, img[..., :nimg], feat_drop_ratio, "img"), img[..., nimg:]), -1)
        cand = torch.cat((_drop(drop, cand[..., :nimg], feat_drop_ratio, "cand"), cand[..., nimg:]), -1)
    q = _drop(drop, h_tilde_prev, drop_ratio, "h_prev")
    visual, alpha_v = soft_dot_attention(q, img, sd["visual_attn.linear_in.weight"])
    x = torch.cat((act, visual), 1)
    h1, c1 = lstm_cell(x, h_tilde_prev, c0, sd["lstm.weight_ih"], sd["lstm.weight_hh"],
                       sd["lstm.bias_ih"], sd["lstm.bias_hh"])
    h1d = _drop(drop, h1, drop_ratio, "h1")
    h_tilde, alpha_c = soft_dot_attention(h1d, ctx, sd["text_attn.linear_in.weight"],
                                          sd["text_attn.linear_out.weight"], ctx_mask)
    htd = _drop(drop, h_tilde, drop_ratio, "h_tilde")
    logit = torch.einsum("bcf,bf->bc", cand, F.linear(htd, sd["cand_attn.weight"]))
    return logit, (h1, c1), h_tilde, (alpha_v, alpha_c)


def follower_decoder(sd, img, a_prev, cands, h0, c0, ctx, ctx_mask, *,
                     drop_ratio=0.5, drop=None):
    """AttnDecoderLSTM.forward, policy.py:37-60."""
    weighted_v, alpha_v = visual_soft_dot_attention(
        h0, img, sd["visual_attn.linear_in_h.weight"], sd["visual_attn.linear_in_h.bias"],
        sd["visual_attn.linear_in_v.weight"], sd["visual_attn.linear_in_v.bias"])
    x = _drop(drop, torch.cat((a_prev, weighted_v), 1), drop_ratio, "x")
    h1, c1 = lstm_cell(x, h0, c0, sd["lstm.weight_ih"], sd["lstm.weight_hh"],
                       sd["lstm.bias_ih"], sd["lstm.bias_hh"])
    h1d = _drop(drop, h1, drop_ratio, "h1")
    h_tilde, alpha_c = soft_dot_attention(h1d, ctx, sd["text_attn.linear_in.weight"],
                                          sd["text_attn.linear_out.weight"], ctx_mask)
    logit = action_scoring(sd, "decode_action.", cands, h_tilde)
    return logit, (h1, c1), (alpha_c, alpha_v)


def monitor_decoder(sd, a_prev, cands, h0, c0, ctx, ctx_mask, cand_mask, *,
                    drop_ratio=0.5, training=False, drop=None):
    """MonitorDecoder.forward, policy.py:132-166."""
    B, C, Fdim = cands.shape
    proj_prev = mlp_with_bn(sd, "proj_navigable_mlp.", a_prev, training, drop, "mlp_prev")
    proj_c = mlp_with_bn(sd, "proj_navigable_mlp.", cands.reshape(-1, Fdim), training,
                         drop, "mlp_cand").view(B, C, -1)
    proj_c = proj_c * (1 - cand_mask.float()).unsqueeze(2)
    pos_ctx = ctx + sd["position.pe"][:, :ctx.shape[1]]                 # units.py:205-207
    if training:
        pos_ctx = _drop(drop, pos_ctx, 0.1, "pos")
    w_ctx, ctx_attn = soft_dot_attention(h0, pos_ctx, sd["text_attn.linear_in.weight"],
                                         None, ctx_mask)
    w_cand, cand_attn = visual_soft_dot_attention(
        h0, proj_c, sd["visual_attn.linear_in_h.weight"], sd["visual_attn.linear_in_h.bias"],
        mask=cand_mask)
    x = torch.cat((proj_prev, w_cand, w_ctx), 1)
    h1, c1 = lstm_cell(x, h0, c0, sd["lstm.weight_ih"], sd["lstm.weight_hh"],
                       sd["lstm.bias_ih"], sd["lstm.bias_hh"])
    h1d = _drop(drop, h1, drop_ratio, "h1") if training else h1
    h_t = F.linear(torch.cat((w_ctx, h1d), 1), sd["action_linear.weight"], sd["action_linear.bias"])
    logit = torch.einsum("bcd,bd->bc", proj_c, h_t)                     # policy.py:108-117
    g = F.linear(torch.cat((h0, w_cand), 1), sd["monitor_linear.weight"], sd["monitor_linear.bias"])
    h_pm = torch.sigmoid(g) * torch.tanh(c1)                            # policy.py:119-130
    if training:
        h_pm = _drop(drop, h_pm, drop_ratio, "h_pm")
    prog = torch.tanh(F.linear(torch.cat((ctx_attn, h_pm), 1), sd["critic.0.weight"],
                               sd["critic.0.bias"])).squeeze()
    return (logit, prog), (h1, c1), (ctx_attn, cand_attn)


def critic(sd, state, drop_ratio=0.5, drop=None):
    """Critic.forward, policy.py:263-267."""
    x = torch.relu(F.linear(state, sd["state2value.0.weight"], sd["state2value.0.bias"]))
    x = _drop(drop, x, drop_ratio, "critic")
    return F.linear(x, sd["state2value.3.weight"], sd["state2value.3.bias"]).squeeze()


def positional_encoding(d_model, max_len=80):
    """PositionalEncoding buffer, units.py:195-203."""
    pe = torch.zeros(max_len, d_model)
    pos = torch.arange(0, max_len).float().unsqueeze(1)
    div = torch.exp(torch.arange(0, d_model, 2).float() * -(math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.unsqueeze(0)
