"""Torch-facing wrappers of the C-ABI kernels: a FeatureStore (HBM tables + TMA context) and
torch.autograd.Function adapters.  PyTorch is plumbing here (memory, streams, autograd tape);
the arithmetic runs in csrc/.
"""
import ctypes as C

import torch

from . import _lib

NSLOT = 16
CMAX = 15
F_DIM = 2176
IMG_DIM = 2048
N_VIEWS = 36


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32c(t):
    assert t.is_cuda and t.dtype == torch.float32, "expected a CUDA fp32 tensor"
    return t.contiguous()


def _i32c(t):
    assert t.is_cuda, "expected a CUDA tensor"
    return t.to(torch.int32).contiguous()


class Rng:
    """Philox stream bookkeeping: one seed, a fresh offset per dropout / sampling call site."""

    def __init__(self, seed=2020):
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.offset = 0

    def next(self):
        self.offset += 1
        return self.seed, self.offset


class FeatureStore:
    """The HBM-resident feature table + index tables of one World, and the library context that
    holds the table's TMA descriptor.  Replaces the reference's in-RAM feature dict
    (misc.py:254-279) and per-step numpy marshalling (base.py:141-157)."""

    def __init__(self, tables, device):
        self.device = torch.device(device)
        assert self.device.type == "cuda", "FeatureStore needs a CUDA device (no CPU path)"
        self.t = tables
        self.table = tables["table"]
        assert self.table.dtype == torch.bfloat16 and self.table.is_contiguous()
        self.n_vp = self.table.shape[0]
        h = C.c_void_p()
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        _lib.check(_lib.lib().vln_ctx_create(C.byref(h), _ptr(self.table), self.n_vp, idx), "vln_ctx_create")
        self.handle = h
        for k in ("loc4", "pose4", "cand_vp", "cand_view", "cand_ang4", "n_cand", "next_hop", "dist", "sq_off",
                  "vp_local"):
            setattr(self, k, tables[k])

    @classmethod
    def from_world(cls, world, device):
        return cls(world.device_tables(device), device)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.lib().vln_ctx_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


# ---- bit-exact gathers ---------------------------------------------------------------------------
def gather_pano(store, vp, view):
    vp, view = _i32c(vp), _i32c(view)
    B = vp.shape[0]
    out = torch.empty((B, N_VIEWS, F_DIM), device=vp.device, dtype=torch.float32)
    _lib.check(_lib.lib().vln_gather_pano(store.handle, _ptr(vp), _ptr(view), _ptr(store.loc4), _ptr(out), B,
                                          _stream()), "vln_gather_pano")
    return out


def gather_cand(store, vp, view, C_slots=NSLOT):
    vp, view = _i32c(vp), _i32c(view)
    B = vp.shape[0]
    out = torch.empty((B, C_slots, F_DIM), device=vp.device, dtype=torch.float32)
    lens = torch.empty((B,), device=vp.device, dtype=torch.int32)
    _lib.check(_lib.lib().vln_gather_cand(store.handle, _ptr(vp), _ptr(view), _ptr(store.cand_view),
                                          _ptr(store.cand_ang4), _ptr(store.n_cand), _ptr(out), _ptr(lens), B,
                                          C_slots, _stream()), "vln_gather_cand")
    return out, lens


def pose_feature(store, view):
    """make_angle_feat(heading, elevation) of the agent's pose (envdrop.py:76-78): [B,128]."""
    return store.pose4[view.long()].repeat_interleave(32, dim=1)


# ---- fused gather + panorama attention ---------------------------------------------------------
def pano_attn_raw(store, vp, view, vec, attn, mode, drop_p=0.0, seed=0, offset=0, split=4):
    B = vp.shape[0]
    out = torch.empty((B, F_DIM), device=vec.device, dtype=torch.float32)
    _lib.check(_lib.lib().vln_pano_attn(store.handle, _ptr(vp), _ptr(view), _ptr(store.loc4), _ptr(vec), _ptr(attn),
                                        _ptr(out), B, mode, float(drop_p), seed, offset, split, _stream()),
               "vln_pano_attn")
    return out


class _PanoAttn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, store, vp, view, drop_p, seed, offset, split):
        q = _f32c(q)
        attn = torch.empty((q.shape[0], N_VIEWS), device=q.device, dtype=torch.float32)
        out = pano_attn_raw(store, vp, view, q, attn, 0, drop_p, seed, offset, split)
        ctx.store, ctx.cfg = store, (drop_p, seed, offset, split)
        ctx.save_for_backward(vp, view, attn)
        ctx.mark_non_differentiable(attn)
        return out, attn

    @staticmethod
    def backward(ctx, d_out, _d_attn):
        vp, view, attn = ctx.saved_tensors
        drop_p, seed, offset, split = ctx.cfg
        dq = pano_attn_raw(ctx.store, vp, view, _f32c(d_out), attn, 1, drop_p, seed, offset, split)
        return dq, None, None, None, None, None, None, None


def pano_attn(store, vp, view, q, drop_p=0.0, seed=0, offset=0, split=4):
    """(weighted [B,2176], attn [B,36]) = softmax_v(x~_v . q) over the episode's panorama."""
    return _PanoAttn.apply(q, store, _i32c(vp), _i32c(view), drop_p, seed, offset, split)


# ---- candidate logits -----------------------------------------------------------------------------
class _CandLogits(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tgt, bias, store, vp, view, drop_p, seed, offset):
        tgt = _f32c(tgt)
        bias = _f32c(bias) if bias is not None else None
        B = tgt.shape[0]
        logits = torch.empty((B, NSLOT), device=tgt.device, dtype=torch.float32)
        _lib.check(_lib.lib().vln_cand_logits_fwd(store.handle, _ptr(vp), _ptr(view), _ptr(store.cand_view),
                                                  _ptr(store.cand_ang4), _ptr(store.n_cand), _ptr(tgt), _ptr(bias),
                                                  _ptr(logits), B, float(drop_p), seed, offset, _stream()),
                   "vln_cand_logits_fwd")
        ctx.store, ctx.cfg, ctx.has_bias = store, (drop_p, seed, offset), bias is not None
        ctx.save_for_backward(vp, view)
        return logits

    @staticmethod
    def backward(ctx, d_logits):
        vp, view = ctx.saved_tensors
        drop_p, seed, offset = ctx.cfg
        store = ctx.store
        d_logits = _f32c(d_logits)
        B = d_logits.shape[0]
        d_tgt = torch.empty((B, F_DIM), device=d_logits.device, dtype=torch.float32)
        d_bias = torch.empty((B,), device=d_logits.device, dtype=torch.float32) if ctx.has_bias else None
        _lib.check(_lib.lib().vln_cand_logits_bwd(store.handle, _ptr(vp), _ptr(view), _ptr(store.cand_view),
                                                  _ptr(store.cand_ang4), _ptr(store.n_cand), _ptr(d_logits),
                                                  _ptr(d_tgt), _ptr(d_bias), B, float(drop_p), seed, offset,
                                                  _stream()), "vln_cand_logits_bwd")
        return d_tgt, d_bias, None, None, None, None, None, None


def cand_logits(store, vp, view, tgt, bias=None, drop_p=0.0, seed=0, offset=0):
    """[B,16] masked candidate logits straight from the table (no [B,C,2176] tensor)."""
    return _CandLogits.apply(tgt, bias, store, _i32c(vp), _i32c(view), drop_p, seed, offset)


# ---- instruction-context attention ------------------------------------------------------------------
class _CtxAttn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, context, tgt, lengths):
        context, tgt = _f32c(context), _f32c(tgt)
        B, L, H = context.shape
        attn = torch.empty((B, L), device=tgt.device, dtype=torch.float32)
        weighted = torch.empty((B, H), device=tgt.device, dtype=torch.float32)
        _lib.check(_lib.lib().vln_ctx_attn_fwd(_ptr(context), _ptr(tgt), _ptr(lengths), _ptr(attn), _ptr(weighted),
                                               B, L, H, _stream()), "vln_ctx_attn_fwd")
        ctx.save_for_backward(context, tgt, lengths, attn)
        return weighted, attn

    @staticmethod
    def backward(ctx, d_weighted, d_attn):
        context, tgt, lengths, attn = ctx.saved_tensors
        B, L, H = context.shape
        d_tgt = torch.empty_like(tgt)
        d_context = torch.zeros_like(context)
        d_weighted = _f32c(d_weighted) if d_weighted is not None else torch.zeros_like(tgt)
        d_attn = _f32c(d_attn) if d_attn is not None else None
        _lib.check(_lib.lib().vln_ctx_attn_bwd(_ptr(context), _ptr(tgt), _ptr(lengths), _ptr(attn), _ptr(d_weighted),
                                               _ptr(d_attn), _ptr(d_tgt), _ptr(d_context), B, L, H, _stream()),
                   "vln_ctx_attn_bwd")
        return d_context, d_tgt, None


def ctx_attn(context, tgt, lengths):
    """(weighted [B,H], attn [B,L]): softmax over the first lengths[b] rows of context[b]."""
    return _CtxAttn.apply(context, tgt, _i32c(lengths))


# ---- LSTM pointwise ----------------------------------------------------------------------------------
class _LstmPointwise(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gates, c0):
        gates, c0 = _f32c(gates), _f32c(c0)
        B, H = c0.shape
        h1, c1, acts = torch.empty_like(c0), torch.empty_like(c0), torch.empty_like(gates)
        _lib.check(_lib.lib().vln_lstm_pointwise_fwd(_ptr(gates), _ptr(c0), _ptr(h1), _ptr(c1), _ptr(acts), B, H,
                                                     _stream()), "vln_lstm_pointwise_fwd")
        ctx.save_for_backward(acts, c0, c1)
        return h1, c1

    @staticmethod
    def backward(ctx, d_h1, d_c1):
        acts, c0, c1 = ctx.saved_tensors
        B, H = c0.shape
        d_gates, d_c0 = torch.empty_like(acts), torch.empty_like(c0)
        d_h1 = _f32c(d_h1) if d_h1 is not None else None
        d_c1 = _f32c(d_c1) if d_c1 is not None else None
        _lib.check(_lib.lib().vln_lstm_pointwise_bwd(_ptr(acts), _ptr(c0), _ptr(c1), _ptr(d_h1), _ptr(d_c1),
                                                     _ptr(d_gates), _ptr(d_c0), B, H, _stream()),
                   "vln_lstm_pointwise_bwd")
        return d_gates, d_c0


def lstm_pointwise(gates, c0):
    return _LstmPointwise.apply(gates, c0)


# ---- action head ---------------------------------------------------------------------------------------
FEEDBACK = {"teacher": 0, "argmax": 1, "sample": 2}


class _Policy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, feedback, seed, offset):
        logits = _f32c(logits)
        B = logits.shape[0]
        dev = logits.device
        ce = torch.empty((B,), device=dev, dtype=torch.float32)
        logp, ent = torch.empty_like(ce), torch.empty_like(ce)
        action = torch.empty((B,), device=dev, dtype=torch.int32)
        probs = torch.empty((B, NSLOT), device=dev, dtype=torch.float32)
        _lib.check(_lib.lib().vln_policy_fwd(_ptr(logits), _ptr(target), feedback, seed, offset, _ptr(ce),
                                             _ptr(action), _ptr(logp), _ptr(ent), _ptr(probs), B, _stream()),
                   "vln_policy_fwd")
        ctx.save_for_backward(probs, target, action, ent)
        ctx.mark_non_differentiable(action)
        return ce, logp, ent, action

    @staticmethod
    def backward(ctx, g_ce, g_logp, g_ent, _g_action):
        probs, target, action, ent = ctx.saved_tensors
        B = probs.shape[0]
        d = torch.empty_like(probs)
        g = [_f32c(x) if x is not None else None for x in (g_ce, g_logp, g_ent)]
        _lib.check(_lib.lib().vln_policy_bwd(_ptr(probs), _ptr(target), _ptr(action), _ptr(ent), _ptr(g[0]),
                                             _ptr(g[1]), _ptr(g[2]), _ptr(d), B, _stream()), "vln_policy_bwd")
        return d, None, None, None, None


def policy_head(logits, target, feedback, seed=0, offset=0):
    """logits [B,16] (-inf = masked) -> (ce [B], logp [B], entropy [B], action int32 [B])."""
    if logits.shape[1] != NSLOT:
        pad = logits.new_full((logits.shape[0], NSLOT - logits.shape[1]), float("-inf"))
        logits = torch.cat((logits, pad), 1)
    fb = FEEDBACK[feedback] if isinstance(feedback, str) else int(feedback)
    return _Policy.apply(logits, _i32c(target) if target is not None else None, fb, seed, offset)


# ---- dropout ----------------------------------------------------------------------------------------------
class _Dropout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p, seed, offset):
        x = _f32c(x)
        y = torch.empty_like(x)
        _lib.check(_lib.lib().vln_dropout(_ptr(x), _ptr(y), x.numel(), float(p), seed, offset, _stream()),
                   "vln_dropout")
        ctx.cfg = (p, seed, offset)
        return y

    @staticmethod
    def backward(ctx, g):
        p, seed, offset = ctx.cfg
        g = _f32c(g)
        y = torch.empty_like(g)
        _lib.check(_lib.lib().vln_dropout(_ptr(g), _ptr(y), g.numel(), float(p), seed, offset, _stream()),
                   "vln_dropout")
        return y, None, None, None


def dropout(x, p, seed, offset):
    if p <= 0.0:
        return x
    return _Dropout.apply(x, p, seed, offset)


def dropout_mask(shape, p, seed, offset, device):
    """The keep-mask (uint8) the kernels use for a tensor of this shape under (seed, offset)."""
    m = torch.empty(shape, device=device, dtype=torch.uint8)
    _lib.check(_lib.lib().vln_dropout_mask(_ptr(m), m.numel(), float(p), seed, offset, _stream()),
               "vln_dropout_mask")
    return m


# ---- environment ---------------------------------------------------------------------------------------------
def env_observe(store, vp, ended, goal):
    B = vp.shape[0]
    teacher = torch.empty((B,), device=vp.device, dtype=torch.int32)
    dist = torch.empty((B,), device=vp.device, dtype=torch.float32)
    _lib.check(_lib.lib().vln_env_observe(_ptr(vp), _ptr(ended), _ptr(goal), _ptr(store.cand_vp), _ptr(store.n_cand),
                                          _ptr(store.next_hop), _ptr(store.dist), _ptr(store.sq_off),
                                          _ptr(store.vp_local), _ptr(teacher), _ptr(dist), B, _stream()),
               "vln_env_observe")
    return teacher, dist


def env_step(store, vp, view, ended, goal, action, last_dist):
    """In-place transition of (vp, view, ended, last_dist); returns (teacher, reward, mask)."""
    B = vp.shape[0]
    teacher = torch.empty((B,), device=vp.device, dtype=torch.int32)
    reward = torch.empty((B,), device=vp.device, dtype=torch.float32)
    mask = torch.empty((B,), device=vp.device, dtype=torch.float32)
    _lib.check(_lib.lib().vln_env_step(_ptr(vp), _ptr(view), _ptr(ended), _ptr(goal), _ptr(action),
                                       _ptr(store.cand_vp), _ptr(store.cand_view), _ptr(store.n_cand),
                                       _ptr(store.next_hop), _ptr(store.dist), _ptr(store.sq_off),
                                       _ptr(store.vp_local), _ptr(last_dist), _ptr(teacher), _ptr(reward),
                                       _ptr(mask), B, _stream()), "vln_env_step")
    return teacher, reward, mask
