"""One decoder step per agent family on the sm_100a kernels.

Constructor arguments, ``forward`` signatures, return tuples and ``state_dict`` keys follow the
reference's ``src/model/policy.py`` (AttnDecoderLSTM :15-60, MonitorDecoder :67-166,
EnvDropDecoder :173-246, Critic :249-267).  ``img_feature`` / ``cand_feature`` / ``a_t_cands``
may be ops.PanoView / ops.CandView handles on the HBM feature table; then nothing of shape
[B,36,2176] or [B,C,2176] is ever materialised and the candidate logits come back already
masked with -inf past each episode's END slot (the agent's length2mask + masked_fill_,
envdrop.py:166-173, is folded into the kernel).
"""
import torch
import torch.nn as nn

from .. import ops
from . import units as U


class AttnDecoderLSTM(U.KernelModule):
    """Speaker-Follower decoder (policy.py:15-60)."""

    def __init__(self, hidden_size, drop_ratio, action_embed_size=2048 + 128, feature_size=2048 + 128,
                 image_attn_layers=None):
        super().__init__()
        self.action_embed_size, self.feature_size, self.hidden_size = action_embed_size, feature_size, hidden_size
        self.drop_ratio = drop_ratio
        self.lstm = nn.LSTMCell(action_embed_size + feature_size, hidden_size)
        self.text_attn = U.SoftDotAttention(hidden_size)
        self.visual_attn = U.VisualSoftDotAttention(hidden_size, feature_size)
        self.decode_action = U.ActionScoring(action_embed_size, hidden_size)

    def forward(self, img_feature, a_t_prev, a_t_cands, h_0, c_0, ctx, ctx_mask=None):
        weighted_v, alpha_v = self.visual_attn(h_0, img_feature)
        visual_ctx = self._drop(torch.cat((a_t_prev, weighted_v), 1), self.drop_ratio, "x")
        h_1, c_1 = ops.lstm_cell(visual_ctx, h_0, c_0, self.lstm.weight_ih, self.lstm.weight_hh,
                                 self.lstm.bias_ih, self.lstm.bias_hh)
        h_1_drop = self._drop(h_1, self.drop_ratio, "h1")
        h_tilde, alpha_c = self.text_attn(h_1_drop, ctx, ctx_mask)
        logit = self.decode_action(a_t_cands, h_tilde)
        return logit, (h_1, c_1), (alpha_c, alpha_v)


class MonitorDecoder(U.KernelModule):
    """Self-Monitoring decoder (policy.py:67-166)."""

    def __init__(self, rnn_hidden_size, drop_ratio, max_enc_len, mlp_dims=(128, 1024),
                 action_embed_size=2048 + 128, feature_size=2048 + 128):
        super().__init__()
        self.rnn_hidden_size, self.max_enc_len = rnn_hidden_size, max_enc_len
        self.mlp_dims = list(mlp_dims)
        self.feature_size, self.action_embed_size = feature_size, action_embed_size
        self.img_hidden_size = self.mlp_dims[-1]
        self.drop_ratio = drop_ratio
        self.proj_navigable_mlp = U.MLPwithBN(input_size=action_embed_size, hidden_size=self.mlp_dims, use_bn=True,
                                              dropout=0.5, use_bias=True, relu=True)
        self.position = U.PositionalEncoding(rnn_hidden_size, dropout=0.1, max_len=max_enc_len)
        self.text_attn = U.SoftDotAttention(rnn_hidden_size, context_only=True)
        self.visual_attn = U.VisualSoftDotAttention(rnn_hidden_size, None, self.img_hidden_size)
        self.lstm = nn.LSTMCell(self.img_hidden_size * 2 + rnn_hidden_size, rnn_hidden_size)
        self.action_linear = nn.Linear(rnn_hidden_size * 2, self.img_hidden_size)
        self.monitor_linear = nn.Linear(rnn_hidden_size + self.img_hidden_size, rnn_hidden_size, bias=True)
        self.critic = nn.Sequential(nn.Linear(max_enc_len + rnn_hidden_size, 1), nn.Tanh())

    def policy_net(self, weighted_ctx, hidden, cands_rep):
        h_tilde = ops.linear(torch.cat((weighted_ctx, hidden), 1), self.action_linear.weight,
                             self.action_linear.bias)
        return torch.bmm(cands_rep, h_tilde.unsqueeze(2)).squeeze(2)

    def progress_monitor(self, h_0, c_1, weighted_cands, ctx_attn):
        g = ops.linear(torch.cat((h_0, weighted_cands), 1), self.monitor_linear.weight, self.monitor_linear.bias)
        h_pm = self._drop(torch.sigmoid(g) * torch.tanh(c_1), self.drop_ratio, "h_pm")
        value = torch.tanh(ops.linear(torch.cat((ctx_attn, h_pm), 1), self.critic[0].weight, self.critic[0].bias))
        return value.squeeze()

    def forward(self, img_feature, a_t_prev, a_t_cands, h_0, c_0, ctx, ctx_mask=None, candidate_mask=None):
        # the BatchNorm statistics run over every row of the padded candidate tensor, END and
        # padding rows included (policy.py:144-149), so this decoder needs the materialised rows
        if isinstance(a_t_cands, ops.CandView):
            a_t_cands = a_t_cands.dense()
        B, C, _ = a_t_cands.shape
        cm = candidate_mask.dense() if isinstance(candidate_mask, U.LengthMask) else candidate_mask
        proj_prev = self.proj_navigable_mlp(a_t_prev)
        if isinstance(candidate_mask, U.LengthMask) and self.training:
            # all C slots go through the MLP; its BatchNorm statistics cover the slots j < max(lengths) — the width the
            # reference pads to — taken from the device-side lengths (no host read-back per step)
            width = candidate_mask.lengths.max()
            w = (torch.arange(C, device=a_t_cands.device) < width).to(a_t_cands.dtype).repeat(B)
            proj_cands = self.proj_navigable_mlp(a_t_cands.reshape(-1, self.action_embed_size), row_weight=w,
                                                 n_rows=(width * B).to(a_t_cands.dtype)).view(B, C, -1)
        else:
            proj_cands = self.proj_navigable_mlp(a_t_cands.reshape(-1, self.action_embed_size)).view(B, C, -1)
        proj_cands = proj_cands * (1 - cm.float()).unsqueeze(2)
        positioned = self.position(ctx)
        weighted_ctx, ctx_attn = self.text_attn(h_0, positioned, ctx_mask)
        weighted_cands, cands_attn = self.visual_attn(h_0, proj_cands, candidate_mask)
        x = torch.cat((proj_prev, weighted_cands, weighted_ctx), 1)
        h_1, c_1 = ops.lstm_cell(x, h_0, c_0, self.lstm.weight_ih, self.lstm.weight_hh, self.lstm.bias_ih,
                                 self.lstm.bias_hh)
        logit = self.policy_net(weighted_ctx, self._drop(h_1, self.drop_ratio, "h1"), proj_cands)
        progress = self.progress_monitor(h_0, c_1, weighted_cands, ctx_attn)
        return (logit, progress), (h_1, c_1), (ctx_attn, cands_attn)


class EnvDropDecoder(U.KernelModule):
    """EnvDrop decoder (policy.py:173-246)."""

    def __init__(self, hidden_size, drop_ratio, feat_drop_ratio, action_embed_size=64, angle_feat_size=128,
                 feature_size=2048 + 128):
        super().__init__()
        self.feature_size, self.action_embed_size = feature_size, action_embed_size
        self.angle_feat_size, self.hidden_size = angle_feat_size, hidden_size
        self.drop_ratio, self.feat_drop_ratio = drop_ratio, feat_drop_ratio
        self.act_embed = nn.Sequential(nn.Linear(angle_feat_size, action_embed_size), nn.Tanh())
        self.lstm = nn.LSTMCell(action_embed_size + feature_size, hidden_size)
        self.text_attn = U.SoftDotAttention(hidden_size)
        self.visual_attn = U.SoftDotAttention(hidden_size, context_dim=feature_size, context_only=True)
        self.cand_attn = nn.Linear(hidden_size, feature_size, bias=False)

    def _feature_dropout(self, feat, tag):
        """env_drop on the image dims only (policy.py:226-231).  A table view is annotated with
        the mask's stream and the kernel applies it while loading; a dense tensor is mutated in
        place, as the reference does to its caller's tensors."""
        p = self.feat_drop_ratio if self.training else 0.0
        if isinstance(feat, (ops.PanoView, ops.CandView)):
            feat.drop_p = p
            if p > 0.0:
                feat.rng = self._rng(feat.vp)
                n_img = feat.shape[2] - self.angle_feat_size
                feat.call_off = feat.rng.next(tag, (feat.shape[0], feat.shape[1], n_img), p)
            return feat
        if p > 0.0:
            img = feat[..., :-self.angle_feat_size]
            img.copy_(self._drop(img.contiguous(), p, tag))
        return feat

    def candidate_attn(self, h_tilde_drop, cand_feat):
        target = ops.linear(h_tilde_drop, self.cand_attn.weight)
        if isinstance(cand_feat, ops.CandView):
            return ops.cand_logits(cand_feat.store, cand_feat.vp, cand_feat.view, target, None,
                                   getattr(cand_feat, "drop_p", 0.0), getattr(cand_feat, "rng", None),
                                   getattr(cand_feat, "call_off", 0))
        return torch.bmm(cand_feat, target.unsqueeze(2)).squeeze(2)

    def forward(self, a_t_prev, img_feature, cand_feature, h_tilde_prev, h_0, c_0, ctx, ctx_mask=None,
                already_dropfeat=False):
        act = torch.tanh(ops.linear(a_t_prev, self.act_embed[0].weight, self.act_embed[0].bias))
        prev_act_emb = self._drop(act, self.drop_ratio, "act")
        if not already_dropfeat:
            img_feature = self._feature_dropout(img_feature, "img")
            cand_feature = self._feature_dropout(cand_feature, "cand")
        prev_h1_drop = self._drop(h_tilde_prev, self.drop_ratio, "h_prev")
        visual_feat, _ = self.visual_attn(prev_h1_drop, img_feature)
        x = torch.cat((prev_act_emb, visual_feat), 1)
        # NB the hidden input is h_tilde_prev; the h_0 argument is unused, as in the reference (:238)
        h_1, c_1 = ops.lstm_cell(x, h_tilde_prev, c_0, self.lstm.weight_ih, self.lstm.weight_hh,
                                 self.lstm.bias_ih, self.lstm.bias_hh)
        h_1_drop = self._drop(h_1, self.drop_ratio, "h1")
        h_tilde, _ = self.text_attn(h_1_drop, ctx, ctx_mask)
        h_tilde_drop = self._drop(h_tilde, self.drop_ratio, "h_tilde")
        logit = self.candidate_attn(h_tilde_drop, cand_feature)
        return logit, (h_1, c_1), h_tilde


class Critic(U.KernelModule):
    """A2C value head (policy.py:249-267)."""

    def __init__(self, hidden_size, drop_ratio):
        super().__init__()
        self.hidden_size, self.drop_ratio = hidden_size, drop_ratio
        self.state2value = nn.Sequential(nn.Linear(hidden_size, hidden_size), nn.ReLU(), U.PhiloxDropout(drop_ratio, "critic"),
                                         nn.Linear(hidden_size, 1))

    def forward(self, state):
        x = torch.relu(ops.linear(state, self.state2value[0].weight, self.state2value[0].bias))
        x = self.state2value[2](x)
        return ops.linear(x, self.state2value[3].weight, self.state2value[3].bias).squeeze()
