"""Fused EnvDrop decoder rollout: the T decoder steps of envdrop.py:134-221 as ~11 kernels per step
forward and ~12 backward, with a hand-written backward through time instead of an autograd tape.

The module-level path (model/policy.py: EnvDropDecoder.forward, one autograd node per op) stays the
drop-in for callers that use the nn.Module directly; the agent's rollout goes through here.  Per
step the chain is (reference lines in policy.py):

    q      = W_vin  drop(h~_{t-1})                         :233 + units.py:107   tcgen05 GEMM
    v      = pano_attn(table[vp_t], q)                      :234, units.py:108-118 fused gather+attention
    gates  = [W_ih | W_hh] [act_emb | v | h~_{t-1}] + b     :236-238              ONE GEMM over the concatenation
    h, c   = LSTM pointwise; WH = [. | drop(h)]             :238-240
    tq     = W_tin drop(h)                                  units.py:107
    WH[:H] = ctx_attn(ctx, tq)                              units.py:108-118
    pre    = W_out WH                                       units.py:119-120
    h~_t   = tanh(pre); hc = drop(h~_t)                     :243
    tgt    = W_cand hc                                      :199-206
    logit  = cand_logits(table, tgt)                        :205 + envdrop.py:166-173
    action, ce, logp, entropy = policy head                 envdrop.py:177-195
    state' = env step                                       envdrop.py:198-219
    XH'[:64] = drop(tanh(W_a pose(view') + b_a))            :222-223

The GEMM operands are row-strided buffers, so no torch.cat / elementwise kernel remains.  Backward
walks the steps in reverse with the transposed weight splits; the five weight gradients are ONE
[N x T*B] x [T*B x K] GEMM each at the end (dY and X of every step are kept stacked).
"""
import ctypes as C

import torch

from .. import ops
from ..ops import _call, _ptr, _stream

H_ACT = 64
FUSE_TAIL = [__import__("os").environ.get("VLN_FUSE_TAIL", "1") != "0"]   # candidate logits + policy/env/act as one launch
# LSTM pointwise + text attention as one launch per step, linear_in folded into the context (csrc/ctx_step.cu)
CTX_STEP = [__import__("os").environ.get("VLN_CTX_STEP", "1") != "0"]
# tanh / dropout glue around h~ as a tile epilogue of its producer GEMM (vln_linear_state_fwd / _bwd, csrc/gemm.cu)
# (bit 0: forward state_fwd in the linear_out GEMM; bit 1: backward state_bwd in the visual_attn.linear_in gradient GEMM)
# Default off: on the B = 64 iteration the separate state_fwd / state_bwd launches are FASTER (4.30 ms vs 4.41 ms with
# both epilogues): a kernel boundary under programmatic dependent launch costs ~1 us, the fence + atomic + spin + L2
# read-back of the tile epilogue ~2-4 us.  Kept selectable (VLN_EPI_STATE=1|2|3) and parity-tested.
EPI_STATE = [int(__import__("os").environ.get("VLN_EPI_STATE", "0"))]


def _p(t, off=0):
    """Device pointer of tensor `t` advanced by `off` fp32 elements."""
    return C.c_void_p(t.data_ptr() + 4 * off)


class _CatSplit:
    """bf16 hi/lo split of [W_ih | W_hh] (and of its transpose), refreshed with the weight epoch."""

    def __init__(self):
        self.stamp = None
        self.sw = None
        self.cat = None

    def fresh(self, w_ih, w_hh):
        stamp = (w_ih._version, w_hh._version, ops.WEIGHT_EPOCH[0], w_ih.data_ptr(), w_hh.data_ptr())
        if stamp != self.stamp:
            self.cat = torch.cat((w_ih.detach(), w_hh.detach()), 1).contiguous()
            if self.sw is None:
                self.sw = ops._SplitWeight(self.cat)
            self.sw.stamp = None
            self.sw.fresh(self.cat)
            self.stamp = stamp
        return self.sw


def _gemm(hi, lo, N, K, x, ldx, M, bias, y, ldy, accumulate=1):
    """y[M,N] += x[M,K] W^T (+ bias) on the tcgen05 kernel (y starts zero-filled): one split-K launch for a batch of
    at most 128 rows, one tall launch (128 x 128 output blocks) for the stacked rows of a whole rollout."""
    if M <= 128:
        _call("vln_linear_bf16x3", _ptr(hi), _ptr(lo), N, K, x, ldx, M, bias, y, ldy, accumulate, 0, _stream())
    else:
        _call("vln_linear_bf16x3_tall", _ptr(hi), _ptr(lo), N, K, x, ldx, M, bias, y, ldy, accumulate, _stream())


_EXP = __import__("os").environ.get("VLN_EXP", "")     # timing experiments only (wrong results): "nofill", "nomask"


def _carver(total, dev):
    slab = (torch.empty if "nofill" in _EXP else torch.zeros)(total, device=dev)
    cur = [0]

    def carve(*shape):
        k = 1
        for s_ in shape:
            k *= s_
        t_ = slab[cur[0]:cur[0] + k].view(*shape)
        cur[0] += k
        return t_
    return carve


def _make_buffers(S, T, B, L, H, dev, paired_rollouts):
    """Work buffers of one rollout, forward and backward.  The GEMM outputs (split-K partial sums meet there) come
    out of zero-filled slabs; with paired rollouts the stacked per-step buffers are zero-filled too (rows that sit
    out later steps must read as finite zeros in the stacked weight-gradient GEMMs).  ~0.5 GB of fills per
    iteration at B = 128: FusedDecoder.prepare issues them on the side stream, under the instruction encoder."""
    F, G4, KX = ops.F_DIM, 4 * H, H_ACT + ops.F_DIM + H
    alloc = torch.zeros if (paired_rollouts and "nofill" not in _EXP) else torch.empty
    c = _carver(S * B * (F + G4) + T * B * (H + H + F), dev)
    f = dict(Q=c(S, B, F), GATES=c(S, B, G4), TQ=c(T, B, H), PRE=c(T, B, H), TGT=c(T, B, F),
             XH=alloc((S + 1, B, KX), device=dev), HQ=alloc((S + 1, B, H), device=dev), HC=alloc((T, B, H), device=dev),
             ACT=alloc((S + 1, B, H_ACT), device=dev), ACTS=alloc((S, B, G4), device=dev),
             CS=alloc((S + 1, B, H), device=dev), H1=alloc((S, B, H), device=dev), WH=alloc((T, B, 2 * H), device=dev),
             ATTV=alloc((S, B, ops.N_VIEWS), device=dev), ATTC=alloc((T, B, L), device=dev),
             LOGIT=alloc((T, B, ops.NSLOT), device=dev), PROBS=alloc((T, B, ops.NSLOT), device=dev),
             CE=torch.zeros((T, B), device=dev), LOGP=torch.zeros((T, B), device=dev), ENT=torch.zeros((T, B), device=dev),
             REWARD=torch.zeros((T, B), device=dev), MASK=torch.zeros((T, B), device=dev),
             ACTION=torch.full((T, B), -1, dtype=torch.int32, device=dev),
             TEACH=torch.full((T + 1, B), -1, dtype=torch.int32, device=dev))
    c = _carver(T * B * (H + 2 * H + KX + H), dev)
    b = dict(DHC=c(T, B, H), DWH=c(T, B, 2 * H), DXH=c(T, B, KX), DHQ=c(T, B, H),
             DTGT=alloc((T, B, F), device=dev), DPRE=alloc((T, B, H), device=dev), DTQ=alloc((T, B, H), device=dev),
             DGATES=alloc((T, B, G4), device=dev), DQ=alloc((T, B, F), device=dev), DACT=alloc((T, B, H_ACT), device=dev),
             DC=alloc((2, B, H), device=dev), DLC=alloc((T, B, L), device=dev))
    return f, b


class FusedDecoder:
    """The EnvDrop decoder rollout as one autograd node.  ``run`` returns per-step stacks
    (ce, logp, entropy [n,B]; h_1 [n,B,H]) that carry gradients, plus detached logits / actions /
    teacher targets / rewards / masks and the bootstrap hidden state."""

    def __init__(self, decoder):
        self.dec = decoder
        self.cat = _CatSplit()
        self._side = None
        self._prep = None
        self.async_wgrad = False       # set by TrainStep (which always differentiates with .backward() into .grad buffers)
        self._wgrad_stream = None
        self.chain_flags = None        # counters of the chain links (ops.chain_begin)
        self.counters = None           # arrive / depart counters of the GEMM tile epilogues (zero between launches)
        self.on_grads_ready = None     # set by TrainStep under data parallelism: called (on the stream that accumulated them)
        #                                once this node's weight gradients are final -> early all-reduce of their bucket

    def params(self):
        d = self.dec
        return [d.act_embed[0].weight, d.act_embed[0].bias, d.lstm.weight_ih, d.lstm.weight_hh, d.lstm.bias_ih,
                d.lstm.bias_hh, d.text_attn.linear_in.weight, d.text_attn.linear_out.weight,
                d.visual_attn.linear_in.weight, d.cand_attn.weight]

    def prepare(self, rng, B, T, feedback, bootstrap, device, pair=None, L=None):
        """Before the encoder runs: hand out the dropout / sampling stream offsets of every decoder pass (in
        the call order of EnvDropDecoder.forward) and draw all feature-dropout keep-bits of the rollout with
        ONE kernel on a side stream, so that it overlaps with the instruction encoder."""
        dec = self.dec
        H = dec.hidden_size
        p = dec.drop_ratio if dec.training else 0.0
        pf = dec.feat_drop_ratio if dec.training else 0.0
        fb = ops.FEEDBACK[feedback]
        S = T + (1 if bootstrap else 0)
        if self.counters is None:
            self.counters = torch.zeros(16, dtype=torch.int32, device=device)
        offs = []
        for t in range(S):
            d = dict(act=0, img=0, cand=0, hprev=0, h1=0, ht=0, sample=0)
            if p > 0.0:
                d["act"] = rng.next("act", (B, H_ACT), p)
            if pf > 0.0:
                d["img"] = rng.next("img", (B, ops.N_VIEWS, ops.IMG_DIM), pf)
                d["cand"] = rng.next("cand", (B, ops.NSLOT, ops.IMG_DIM), pf)
            if p > 0.0:
                d["hprev"] = rng.next("h_prev", (B, H), p)
                d["h1"] = rng.next("h1", (B, H), p)
                d["ht"] = rng.next("h_tilde", (B, H), p)
            if fb == 2 and t < T:
                d["sample"] = rng.next()
            offs.append(d)
        MB, side, bufs = None, None, None
        if pf > 0.0 or L is not None:
            main = torch.cuda.current_stream()
            if self._side is None:
                self._side = torch.cuda.Stream()
            side = self._side
            side.wait_stream(main)      # also orders this iteration's fills after every earlier use of recycled blocks
        # Feature-dropout keep-bits (policy.py:226-231) of every decoder pass, packed, drawn by one kernel on the side
        # stream.  (The panorama kernel can also draw a launch's bits itself ahead of its dependency wait — what it does
        # for callers without pre-generated bits — but on the step chain that work is not hidden: it shares the SM with
        # the predecessor's CTAs and slowed the chain by as much as the bulk kernel costs; measured 4.35 vs 4.37 ms.)
        n_bits = S if pf > 0.0 else 0
        if n_bits > 0:
            MB = torch.empty((n_bits, B * ops.N_VIEWS, ops.IMG_DIM // 8), dtype=torch.uint8, device=device)   # owned by `main`
        if n_bits > 0 and "nomask" not in _EXP:
            with torch.cuda.stream(side):
                stride = offs[1]["img"] - offs[0]["img"] if S > 1 else 0
                V = ops.N_VIEWS
                if pair is None or n_bits < S:
                    _call("vln_feature_mask_bits", _ptr(MB), B * V, n_bits, pf, rng.ptr, offs[0]["img"], stride, _stream())
                else:        # rows [0, B_main) for every pass, the teacher-forced rows only for their T_teacher steps
                    B_main, T_t = pair
                    _call("vln_feature_mask_bits_ld", _ptr(MB), B_main * V, B * V, 0, S, pf, rng.ptr, offs[0]["img"], stride,
                          _stream())
                    _call("vln_feature_mask_bits_ld", _ptr(MB), (B - B_main) * V, B * V, B_main * V, T_t, pf, rng.ptr,
                          offs[0]["img"], stride, _stream())
        if side is not None:            # this iteration's bf16 weight splits (they only depend on the weights): off the
            with torch.cuda.stream(side):        # chain between encoder and decoder
                self.cat.fresh(dec.lstm.weight_ih, dec.lstm.weight_hh)
                for w in self.params()[6:10]:
                    ops._split_of(w)
        if L is not None:               # the rollout's work buffers: allocated + zero-filled under the encoder
            with torch.cuda.stream(side):
                bufs = _make_buffers(S, T, B, L, dec.hidden_size, device, pair is not None)
                for d_ in bufs:
                    for t_ in d_.values():
                        t_.record_stream(main)
        return dict(offs=offs, MB=MB, side=side, p=p, pf=pf, S=S, bufs=bufs, shape=(S, T, B, L))

    def run(self, rng, st, ctx, lengths, h0, c0, T, feedback, bootstrap, poll, split, prep=None, pair=None):
        """``pair = (B_main, T_teacher)``: the batch holds two rollouts of the same minibatch stepped together
        (trainer.py:411-421) — rows [0, B_main) follow ``feedback`` for T steps, rows [B_main, B) are
        teacher-forced and only take part in the first T_teacher steps (every launch of a later step covers the
        row prefix [0, B_main) only)."""
        assert pair is None or (poll == 0 and 0 < pair[0] < st.B and 0 < pair[1] <= T)
        if prep is None:
            prep = self.prepare(rng, st.B, T, feedback, bootstrap, ctx.device)
        if prep["side"] is not None:
            torch.cuda.current_stream().wait_stream(prep["side"])
            prep["side"] = None
        self._prep = prep
        return _Rollout.apply(self, rng, st, lengths, T, feedback, bootstrap, poll, split, pair, ctx, h0, c0, *self.params())


N_META = 10         # non-tensor arguments of _Rollout.apply before (ctx, h0, c0, *params)


class _Rollout(torch.autograd.Function):
    @staticmethod
    def forward(fctx, fd, rng, st, lengths, T, feedback, bootstrap, poll, split, pair, ctx, h0, c0, *params):
        dec = fd.dec
        w_act, b_act, w_ih, w_hh, b_ih, b_hh = [q.detach() for q in params[:6]]
        ctx, h0, c0 = ops._f32c(ctx.detach()), ops._f32c(h0.detach()), ops._f32c(c0.detach())
        store = st.store
        dev = ctx.device
        B, L, H = ctx.shape
        F, G4, KX = ops.F_DIM, 4 * H, H_ACT + ops.F_DIM + H
        p = dec.drop_ratio if dec.training else 0.0
        pf = dec.feat_drop_ratio if dec.training else 0.0
        fb = ops.FEEDBACK[feedback]
        S = T + (1 if bootstrap else 0)                       # decoder passes (the bootstrap one only up to h_1)
        rp = rng.ptr
        B_all = B
        B_main, T_pair = pair if pair is not None else (B, S + 1)
        if pair is not None:
            fb |= (B_main + 1) << 8                           # rows >= B_main are teacher-forced (vln_policy_env_act_fwd)

        def rows(t):                                          # episodes that take part in decoder pass t
            return B_all if t < T_pair else B_main
        # ---- weights: bf16 hi/lo splits, refreshed once per optimiser step ----
        s_cat = fd.cat.fresh(w_ih, w_hh)
        s_tin, s_out, s_vin, s_cand = (ops._split_of(w) for w in params[6:10])
        bsum = (b_ih + b_hh).contiguous()
        w_act = w_act.view(H_ACT, 4, 32).sum(2).contiguous()        # group sums: the angle feature is 4 values x32

        # ---- buffers (zero-filled GEMM output slabs, stacked per-step tensors): prepared under the encoder if possible ----
        prep = fd._prep
        if prep.get("bufs") is not None and prep["shape"] == (S, T, B, L):
            fb_, bb_ = prep["bufs"]
        else:
            fb_, bb_ = _make_buffers(S, T, B, L, H, dev, pair is not None)
        prep["bufs"] = None
        Q, GATES, TQ, PRE, TGT = fb_["Q"], fb_["GATES"], fb_["TQ"], fb_["PRE"], fb_["TGT"]
        XH, HQ, HC, ACT, ACTS, CS, H1, WH = (fb_[k] for k in ("XH", "HQ", "HC", "ACT", "ACTS", "CS", "H1", "WH"))
        ATTV, ATTC, LOGIT, PROBS = fb_["ATTV"], fb_["ATTC"], fb_["LOGIT"], fb_["PROBS"]
        CE, LOGP, ENT, REWARD, MASK, ACTION, TEACH = (fb_[k] for k in ("CE", "LOGP", "ENT", "REWARD", "MASK", "ACTION", "TEACH"))
        TEACH[0].copy_(st.teacher)
        CS[0].copy_(c0)

        offs, MB = prep["offs"], prep["MB"]
        assert prep["S"] == S and prep["p"] == p and prep["pf"] == pf

        def act_embed(t):
            _call("vln_envdrop_act_fwd", _ptr(st.view[t]), _ptr(store.pose4), _ptr(w_act), _ptr(b_act), _ptr(ACT[t]),
                  _ptr(XH[t]), KX, rows(t), H_ACT, p, rp, offs[t]["act"], _stream())

        # text-attention stage as one launch (csrc/ctx_step.cu): CW = ctx W_in once per rollout replaces the per-step
        # query projection tq = W_in drop(h_1)  (logit_l = ctx_l . tq = CW_l . drop(h_1))
        use_cs = CTX_STEP[0] and H == 512 and L <= 80
        use_epi = EPI_STATE[0] if (B <= 128 and fd.counters is not None) else 0
        CW = ops._tc_matmul_tall(ctx.view(B * L, H), s_tin.hi_t, s_tin.lo_t, H, H).view(B, L, H) if use_cs else None

        def visual_and_lstm(t, need_drop, q_done=False, pointwise=True):
            Bt = rows(t)
            if not q_done:
                _gemm(s_vin.hi, s_vin.lo, F, H, _p(HQ[t]), H, Bt, None, _p(Q[t]), F)
            _call("vln_pano_attn_ld", store.handle, _ptr(st.vp[t]), _ptr(st.view[t]), _ptr(store.loc4), _ptr(Q[t]), F,
                  _ptr(ATTV[t]), None, F, _p(XH[t], H_ACT), KX, Bt, 0, pf, rp, offs[t]["img"],
                  _ptr(MB[t]) if (MB is not None and t < MB.shape[0]) else None, split, _stream())
            _gemm(s_cat.hi, s_cat.lo, G4, KX, _p(XH[t]), KX, Bt, _ptr(bsum), _p(GATES[t]), G4)
            if not pointwise:
                return
            _call("vln_lstm_pointwise_drop_fwd", _ptr(GATES[t]), _ptr(CS[t]), _ptr(H1[t]), _ptr(CS[t + 1]),
                  _ptr(ACTS[t]), _p(WH[t], H) if need_drop else None, 2 * H, Bt, H, p, rp, offs[t]["h1"], _stream())

        # from here to the end of the rollout every launch is a kernel of this package: their grid-to-grid dependencies
        # are resolved through device-side counters (chain links, csrc/common.cuh) instead of grid completion
        ops.chain_begin(fd)
        _call("vln_envdrop_state_fwd", _ptr(h0), 0, _p(XH[0], H_ACT + F), KX, _ptr(HQ[0]), None, B, H, p, rp,
              offs[0]["hprev"], 0, _stream())
        act_embed(0)
        n = 0
        paired = B <= 128
        for t in range(T):
            Bt = rows(t)
            visual_and_lstm(t, True, q_done=paired and t > 0, pointwise=not use_cs)
            if use_cs:
                _call("vln_envdrop_ctx_step_fwd", _ptr(GATES[t]), _ptr(CS[t]), _ptr(H1[t]), _ptr(CS[t + 1]), _ptr(ACTS[t]),
                      _ptr(WH[t]), 2 * H, _ptr(ctx), _ptr(CW), _ptr(lengths), _ptr(ATTC[t]), Bt, L, H, p, rp,
                      offs[t]["h1"], 1 if t > 0 else 0, _stream())
            else:
                _gemm(s_tin.hi, s_tin.lo, H, H, _p(WH[t], H), 2 * H, Bt, None, _p(TQ[t]), H)
                _call("vln_ctx_attn_fwd_ld", _ptr(ctx), _ptr(TQ[t]), _ptr(lengths), _ptr(ATTC[t]), _ptr(WH[t]), 2 * H, Bt,
                      L, H, 1 if t > 0 else 0, _stream())
            more = t + 1 < S
            if use_epi & 1:     # pre = W_out [weighted | drop(h)], then h~ = tanh(pre) and its two dropout sites as the tile epilogue
                _call("vln_linear_state_fwd", _ptr(s_out.hi), _ptr(s_out.lo), H, 2 * H, _p(WH[t]), 2 * H, Bt, _p(PRE[t]), H,
                      _p(XH[t + 1], H_ACT + F), KX, _ptr(HQ[t + 1]) if more else None, _ptr(HC[t]), p, rp,
                      offs[t + 1]["hprev"] if more else 0, offs[t]["ht"], _ptr(fd.counters), _stream())
            else:
                _gemm(s_out.hi, s_out.lo, H, 2 * H, _p(WH[t]), 2 * H, Bt, None, _p(PRE[t]), H)
                _call("vln_envdrop_state_fwd", _ptr(PRE[t]), 1, _p(XH[t + 1], H_ACT + F), KX,
                      _ptr(HQ[t + 1]) if more else None, _ptr(HC[t]), Bt, H, p, rp,
                      offs[t + 1]["hprev"] if more else 0, offs[t]["ht"], _stream())
            if paired and more:      # tgt_t = W_cand hc_t and q_{t+1} = W_vin hq_{t+1} both only wait for h~_t: one launch
                _call("vln_linear_bf16x3_pair", _ptr(s_cand.hi), _ptr(s_cand.lo), _ptr(HC[t]), _ptr(TGT[t]), _ptr(s_vin.hi),
                      _ptr(s_vin.lo), _ptr(HQ[t + 1]), _ptr(Q[t + 1]), F, H, H, Bt, F, _stream())
            else:
                _gemm(s_cand.hi, s_cand.lo, F, H, _p(HC[t]), H, Bt, None, _p(TGT[t]), F)
            if FUSE_TAIL[0]:
                # candidate logits + action head + simulator transition + the next pass's action embedding: one launch
                _call("vln_cand_policy_env_act_fwd", store.handle, _ptr(st.vp[t]), _ptr(st.view[t]), _ptr(store.cand_ang4),
                      _ptr(TGT[t]), _ptr(LOGIT[t]), pf, offs[t]["cand"], _ptr(TEACH[t]), fb, rp, offs[t]["sample"],
                      _ptr(CE[t]), _ptr(ACTION[t]), _ptr(LOGP[t]), _ptr(ENT[t]), _ptr(PROBS[t]),
                      _ptr(st.ended[t]), _ptr(st.dist[t]), _ptr(st.goal),
                      _ptr(store.cand_vp), _ptr(store.cand_view), _ptr(store.n_cand), _ptr(store.next_hop),
                      _ptr(store.dist), _ptr(store.sq_off), _ptr(store.vp_local),
                      _ptr(st.vp[t + 1]), _ptr(st.view[t + 1]), _ptr(st.ended[t + 1]), _ptr(st.dist[t + 1]),
                      _ptr(TEACH[t + 1]), _ptr(REWARD[t]), _ptr(MASK[t]), _ptr(st.n_active[t:t + 1]),
                      _ptr(store.pose4), _ptr(w_act), _ptr(b_act), _ptr(ACT[t + 1]) if more else None,
                      _ptr(XH[t + 1]) if more else None, KX, H_ACT, p, offs[t + 1]["act"] if more else 0, Bt, _stream())
            else:
                _call("vln_cand_logits_fwd", store.handle, _ptr(st.vp[t]), _ptr(st.view[t]), _ptr(store.cand_view),
                      _ptr(store.cand_ang4), _ptr(store.n_cand), _ptr(TGT[t]), None, _ptr(LOGIT[t]), Bt, pf, rp,
                      offs[t]["cand"], _stream())
                # action head + simulator transition + the next pass's action embedding: one launch
                _call("vln_policy_env_act_fwd", _ptr(LOGIT[t]), _ptr(TEACH[t]), fb, rp, offs[t]["sample"], _ptr(CE[t]),
                      _ptr(ACTION[t]), _ptr(LOGP[t]), _ptr(ENT[t]), _ptr(PROBS[t]),
                      _ptr(st.vp[t]), _ptr(st.view[t]), _ptr(st.ended[t]), _ptr(st.dist[t]), _ptr(st.goal),
                      _ptr(store.cand_vp), _ptr(store.cand_view), _ptr(store.n_cand), _ptr(store.next_hop),
                      _ptr(store.dist), _ptr(store.sq_off), _ptr(store.vp_local),
                      _ptr(st.vp[t + 1]), _ptr(st.view[t + 1]), _ptr(st.ended[t + 1]), _ptr(st.dist[t + 1]),
                      _ptr(TEACH[t + 1]), _ptr(REWARD[t]), _ptr(MASK[t]), _ptr(st.n_active[t:t + 1]),
                      _ptr(store.pose4), _ptr(w_act), _ptr(b_act), _ptr(ACT[t + 1]) if more else None,
                      _ptr(XH[t + 1]) if more else None, KX, H_ACT, p, offs[t + 1]["act"] if more else 0, Bt, _stream())
            st.steps = n = t + 1
            if poll and (t + 1) % poll == 0 and t + 1 < T and st.all_ended(t):
                break
        if bootstrap:                                           # envdrop.py:225-237: h_1 of the state after the last step
            visual_and_lstm(n, False, q_done=paired and n > 0)
        ops.chain_end()
        st.teacher = TEACH[n]
        # kept alive for inspection (tests read a CUDA-graph replay's logits / actions from these static buffers)
        fd.last = dict(LOGIT=LOGIT, ACTION=ACTION, TEACH=TEACH, n=n)

        fctx.fd, fctx.st, fctx.rp, fctx.MB, fctx.bwd_bufs = fd, st, rp, MB, bb_
        fctx.cfg = (n, B, L, H, p, pf, split, offs, B_main, T_pair, use_cs, use_epi)
        fctx.splits = (s_cat, s_vin, s_tin, s_out, s_cand)
        fctx.save_for_backward(ctx, lengths, XH, HQ, HC, ACT, ACTS, CS, WH, ATTV, ATTC, CW if use_cs else TQ, PROBS, ENT,
                               ACTION, TEACH)
        outs = (CE[:n], LOGP[:n], ENT[:n], H1[:n], LOGIT[:n], ACTION[:n], TEACH[:n], REWARD[:n], MASK[:n],
                H1[n].clone() if bootstrap else H1[:0].clone())
        fctx.mark_non_differentiable(*outs[4:])
        return outs

    @staticmethod
    def backward(fctx, d_ce, d_logp, d_ent, d_h1, *_unused):
        ctx, lengths, XH, HQ, HC, ACT, ACTS, CS, WH, ATTV, ATTC, TQ, PROBS, ENT, ACTION, TEACH = fctx.saved_tensors
        n, B, L, H, p, pf, split, offs, B_main, T_pair, use_cs, use_epi = fctx.cfg
        CW = TQ                                                 # (the saved slot holds CW = ctx W_in with use_cs)

        def rows(t):
            return B if t < T_pair else B_main
        s_cat, s_vin, s_tin, s_out, s_cand = fctx.splits
        st, rp, MB, fd = fctx.st, fctx.rp, fctx.MB, fctx.fd
        store = st.store
        dev = ctx.device
        F, G4, KX = ops.F_DIM, 4 * H, H_ACT + ops.F_DIM + H
        OH = H_ACT + F                                          # column of h~ inside an XH row
        d_ce, d_logp, d_ent, d_h1 = (ops._f32c(g) if g is not None else None for g in (d_ce, d_logp, d_ent, d_h1))

        bb_ = fctx.bwd_bufs                                       # zero-filled in prepare(): [T, B, ...], the first n steps used
        DHC, DWH, DXH, DHQ = bb_["DHC"][:n], bb_["DWH"][:n], bb_["DXH"][:n], bb_["DHQ"][:n]
        DTGT, DPRE, DTQ, DGATES = bb_["DTGT"][:n], bb_["DPRE"][:n], bb_["DTQ"][:n], bb_["DGATES"][:n]
        DQ, DACT, DC, DLC = bb_["DQ"][:n], bb_["DACT"][:n], bb_["DC"], bb_["DLC"][:n]

        # ---- off the recursion: candidate-logit backward of ALL steps in one launch, then d(h~_drop) = dtgt W_cand
        #      as a stack of 128-row GEMMs (none of it depends on the backward-in-time chain) ----
        stride = (offs[1]["cand"] - offs[0]["cand"]) if len(offs) > 1 else 0
        ops.chain_begin(fd)
        _call("vln_cand_logits_bwd_policy", store.handle, _ptr(st.vp), _ptr(st.view), _ptr(store.cand_view),
              _ptr(store.cand_ang4), _ptr(store.n_cand), _ptr(PROBS), _ptr(TEACH), _ptr(ACTION), _ptr(ENT),
              _ptr(d_ce) if d_ce is not None else None, _ptr(d_logp) if d_logp is not None else None,
              _ptr(d_ent) if d_ent is not None else None, _ptr(DTGT), B, n, pf, rp, offs[0]["cand"], stride, _stream())
        _gemm(s_cand.hi_t, s_cand.lo_t, H, F, _p(DTGT), F, n * B, None, _p(DHC), H)
        for t in range(n - 1, -1, -1):
            last = t == n - 1
            Bt = rows(t)            # rows that sat out step t+1 see zero carried gradients (zero-filled slab / DC)
            if last or not (use_epi & 2) or rows(t + 1) != Bt:   # (else DPRE[t] came out of the epilogue of step t+1's last GEMM)
                _call("vln_envdrop_state_bwd", _ptr(DHC[t]), None if last else _p(DXH[t + 1], OH), KX,
                      None if last else _ptr(DHQ[t + 1]), _p(XH[t + 1], OH), KX, 1, _ptr(DPRE[t]), Bt, H, p, rp,
                      0 if last else offs[t + 1]["hprev"], offs[t]["ht"], _stream())
            _gemm(s_out.hi_t, s_out.lo_t, 2 * H, H, _p(DPRE[t]), H, Bt, None, _p(DWH[t]), 2 * H)
            if use_cs:      # text attention backward + d drop(h_1) += sum_l dlogit_l CW_l + LSTM pointwise backward: one launch
                _call("vln_envdrop_ctx_step_bwd", _ptr(ctx), _ptr(CW), _ptr(lengths), _ptr(ATTC[t]), _ptr(DWH[t]), 2 * H,
                      _ptr(DLC[t]), _ptr(ACTS[t]), _ptr(CS[t]), _ptr(CS[t + 1]),
                      _ptr(d_h1[t]) if d_h1 is not None else None, None if last else _ptr(DC[(t + 1) & 1]),
                      _ptr(DGATES[t]), _ptr(DC[t & 1]), Bt, L, H, p, rp, offs[t]["h1"], _stream())
            else:
                _call("vln_ctx_attn_bwd_ld", _ptr(ctx), _ptr(TQ[t]), _ptr(lengths), _ptr(ATTC[t]), _ptr(DWH[t]), 2 * H, None,
                      _ptr(DTQ[t]), None, _ptr(DLC[t]), Bt, L, H, 1, _stream())
                _gemm(s_tin.hi_t, s_tin.lo_t, H, H, _p(DTQ[t]), H, Bt, None, _p(DWH[t], H), 2 * H, accumulate=1)
                _call("vln_lstm_pointwise_drop_bwd", _ptr(ACTS[t]), _ptr(CS[t]), _ptr(CS[t + 1]), _p(DWH[t], H), 2 * H,
                      _ptr(d_h1[t]) if d_h1 is not None else None, None if last else _ptr(DC[(t + 1) & 1]),
                      _ptr(DGATES[t]), _ptr(DC[t & 1]), Bt, H, p, rp, offs[t]["h1"], _stream())
            _gemm(s_cat.hi_t, s_cat.lo_t, KX, G4, _p(DGATES[t]), G4, Bt, None, _p(DXH[t]), KX)
            _call("vln_pano_attn_ld", store.handle, _ptr(st.vp[t]), _ptr(st.view[t]), _ptr(store.loc4),
                  _p(DXH[t], H_ACT), KX, _ptr(ATTV[t]), _p(XH[t], H_ACT), KX, _ptr(DQ[t]), F, Bt, 1 | 2, pf, rp,
                  offs[t]["img"], _ptr(MB[t]) if (MB is not None and t < MB.shape[0]) else None, split, _stream())
            if (use_epi & 2) and t > 0 and rows(t - 1) == Bt:
                # d_hq_t = dq_t W_vin, then the gradient of h~_{t-1} through tanh and its two dropout sites: the tile epilogue
                _call("vln_linear_state_bwd", _ptr(s_vin.hi_t), _ptr(s_vin.lo_t), H, F, _p(DQ[t]), F, Bt, _p(DHQ[t]), H,
                      _ptr(DHC[t - 1]), _p(DXH[t], OH), KX, _p(XH[t], OH), KX, _ptr(DPRE[t - 1]), p, rp,
                      offs[t]["hprev"], offs[t - 1]["ht"], _ptr(fd.counters), _stream())
            else:
                _gemm(s_vin.hi_t, s_vin.lo_t, H, F, _p(DQ[t]), F, Bt, None, _p(DHQ[t]), H)
        _call("vln_envdrop_act_bwd", _ptr(DXH), KX, _ptr(ACT), _ptr(DACT), B, H_ACT, n, p, rp, offs[0]["act"],
              (offs[1]["act"] - offs[0]["act"]) if len(offs) > 1 else 0, _stream())
        d_h0 = torch.empty((B, H), device=dev)
        _call("vln_envdrop_state_bwd", None, _p(DXH[0], OH), KX, _ptr(DHQ[0]), None, 0, 0, _ptr(d_h0), B, H, p, rp,
              offs[0]["hprev"], 0, _stream())
        ops.chain_end()
        d_c0 = DC[0]

        # ---- d_ctx[b] = sum_t attn_t^T d_weighted_t + dlogit_t^T tq_t: one batched GEMM pair over all steps ----
        d_ctx = ops.seq_outer_sum(ATTC[:n], DWH[:n, :, :H])
        dCW = None
        if use_cs:          # through CW = ctx W_in:  dCW = sum_t dlogit_t (x) drop(h_1)_t ,  d_ctx += dCW W_in^T
            dCW = ops.seq_outer_sum(DLC[:n], WH[:n, :, H:])
            _gemm(s_tin.hi, s_tin.lo, H, H, _p(dCW), H, B * L, None, _p(d_ctx), H, accumulate=1)
        else:
            ops.seq_outer_sum(DLC[:n], TQ[:n], out=d_ctx)

        # ---- weight gradients: one GEMM per weight over all n*B rows ----
        def weight_grads():
            nb = n * B
            dG = DGATES.view(nb, G4)
            d_wcat = ops.wgrad(dG, XH[:n].reshape(nb, KX))
            d_b = dG.sum(0)
            if use_cs:      # dW_in[k, j] = sum_{b,l} ctx[b,l,k] dCW[b,l,j]
                d_w_tin = ops.wgrad(ctx.reshape(B * L, H), dCW.view(B * L, H))
            else:
                d_w_tin = ops.wgrad(DTQ.view(nb, H), WH[:n, :, H:].reshape(nb, H))
            d_w_out = ops.wgrad(DPRE.view(nb, H), WH[:n].reshape(nb, 2 * H))
            d_w_vin = ops.wgrad(DQ.view(nb, F), HQ[:n].reshape(nb, H))
            d_w_cand = ops.wgrad(DTGT.view(nb, F), HC[:n].reshape(nb, H))
            pose = store.pose128[st.view[:n].reshape(-1).long()]
            dA = DACT.view(nb, H_ACT)
            d_w_act = ops.wgrad(dA, pose)
            d_b_act = dA.sum(0)
            return (d_w_act, d_b_act, d_wcat[:, :OH], d_wcat[:, OH:], d_b, d_b, d_w_tin, d_w_out, d_w_vin, d_w_cand)

        fd = fctx.fd
        params = fd.params()
        if fd.async_wgrad and all(q.grad is not None for q in params):
            # Nothing downstream of this node needs the weight gradients (the encoder's backward only takes d_ctx /
            # d_h0 / d_c0), so their GEMMs run on a side stream underneath the encoder's latency-bound BPTT kernel and
            # accumulate straight into the flat gradient buffer; the calling stream joins when the backward pass ends.
            main = torch.cuda.current_stream()
            if fd._wgrad_stream is None:
                fd._wgrad_stream = torch.cuda.Stream()
            side = fd._wgrad_stream
            side.wait_stream(main)
            with torch.cuda.stream(side):
                grads = weight_grads()
                for q, g_ in zip(params, grads):
                    q.grad.add_(g_)
                if fd.on_grads_ready is not None:
                    fd.on_grads_ready()
            keep = [grads, DGATES, XH, WH, HQ, HC, DTQ, DPRE, DQ, DTGT, DACT, dCW, ctx]    # alive until the join (allocator safety)

            def join():
                torch.cuda.current_stream().wait_stream(side)
                keep.clear()
            torch.autograd.Variable._execution_engine.queue_callback(join)
            return (None,) * N_META + (d_ctx, d_h0, d_c0) + (None,) * 10
        return (None,) * N_META + (d_ctx, d_h0, d_c0) + weight_grads()
