"""Torch-facing wrappers of the C-ABI kernels: a FeatureStore (HBM tables + TMA context), lazy
feature views, and torch.autograd.Function adapters.  PyTorch is plumbing here (memory, streams,
autograd tape); the arithmetic runs in csrc/.  There is no CPU path: every op needs CUDA tensors
and the built library.
"""
import ctypes as C

import torch
import torch.nn.functional as F

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
    return t if (t.dtype == torch.int32 and t.is_contiguous()) else t.to(torch.int32).contiguous()


CALLS = [0]       # number of C-ABI kernel-launching calls made (graph replays add their captured count)


def _call(name, *args):
    CALLS[0] += 1
    _lib.check(getattr(_lib.lib(), name)(*args), name)


CHAIN_FLAGS_N = 4096


def chain_begin(owner):
    """Open a chain region on the current stream (include/vln_b200.h: vln_chain_begin): until chain_end() consecutive
    kernels of this package hand over through device-side counters kept in ``owner.chain_flags``."""
    if getattr(owner, "chain_flags", None) is None:
        owner.chain_flags = torch.zeros(CHAIN_FLAGS_N, dtype=torch.int32, device=torch.cuda.current_device())
    _call("vln_chain_begin", _ptr(owner.chain_flags), CHAIN_FLAGS_N, _stream())


def chain_end():
    return _lib.lib().vln_chain_end()


def chain_timeouts(owner):
    """Polls that gave up (must be 0: a non-zero count means a producer signalled fewer arrivals than announced)."""
    f = getattr(owner, "chain_flags", None)
    return 0 if f is None else int(f[-1])


class Rng:
    """Philox stream bookkeeping.  ``state`` is a device int64[2] = {seed, base}; every dropout /
    sampling call site takes the next ``call_off`` and its kernel draws from stream
    (seed, base + call_off).  ``advance()`` moves the base on the device, so a captured CUDA graph
    gets fresh masks on every replay.  ``log`` (when a list) records (tag, shape, p, call_off) of
    every dropout site so tests can regenerate the very same masks for the oracle."""

    def __init__(self, seed=2020, device="cuda"):
        # words 0, 1 = {seed, base} (the C-ABI's rng state); the room behind them is only touched by the debug build's
        # chain stamps (csrc/common.cuh CHAIN_BEGIN: word 2 = enable, word 3 = count, 4 words per record from word 4)
        self.state = torch.zeros(4 + 4 * 8192, dtype=torch.int64, device=device)
        self.state[0] = int(seed) & 0x7FFFFFFFFFFFFFFF
        self.off = 0
        self.log = None

    def next(self, tag=None, shape=None, p=None):
        self.off += 1
        if self.log is not None and tag is not None:
            self.log.append((tag, tuple(shape), float(p), self.off))
        return self.off

    def advance(self):
        """Consume the offsets handed out since the last advance (device-side base += n)."""
        n, self.off = self.off, 0
        if n:
            _call("vln_rng_advance", _ptr(self.state), n, _stream())

    def begin_iteration(self):
        """Restart call-site numbering (the same sites get the same call_off every iteration — a
        requirement for graph replay) after moving the base past the previous iteration."""
        self.advance()

    @property
    def ptr(self):
        return _ptr(self.state)


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
        _call("vln_ctx_create", C.byref(h), _ptr(self.table), self.n_vp, idx)
        self.handle = h
        for k in ("loc4", "pose4", "cand_vp", "cand_view", "cand_ang4", "n_cand", "next_hop", "dist", "sq_off",
                  "vp_local"):
            setattr(self, k, tables[k])
        self.pose128 = self.pose4.repeat_interleave(32, dim=1).contiguous()      # [36,128]

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


class PanoView:
    """Stands for the reference's ``img_feature`` tensor [B,36,2176] (base.py:141-147) without
    materialising it: the rows live in the HBM table, keyed by (viewpoint, current view)."""

    def __init__(self, store, vp, view):
        self.store, self.vp, self.view = store, _i32c(vp), _i32c(view)
        self.shape = (self.vp.shape[0], N_VIEWS, F_DIM)

    def dense(self):
        return gather_pano(self.store, self.vp, self.view)


class CandView:
    """Stands for ``candidate_feat`` [B,C,2176] + ``candidate_leng`` (base.py:149-157)."""

    def __init__(self, store, vp, view):
        self.store, self.vp, self.view = store, _i32c(vp), _i32c(view)
        self.shape = (self.vp.shape[0], NSLOT, F_DIM)

    def dense(self):
        return gather_cand(self.store, self.vp, self.view)[0]

    def lengths(self):
        return self.store.n_cand[self.vp.long()] + 1


# ---- GEMM-shaped pieces ---------------------------------------------------------------------------
WEIGHT_EPOCH = [0]        # bumped by the fused optimiser (it updates parameters through raw pointers)
USE_TC_LINEAR = [True]    # tcgen05 bf16x3 path for skinny linears; False = cuBLAS fp32 everywhere
_split_cache = {}


class _SplitWeight:
    """bf16 hi/lo split of an fp32 weight [N,K] and of its transpose, refreshed lazily whenever the
    parameter changed (autograd version counter, or the optimiser epoch)."""

    def __init__(self, w):
        N, K = w.shape
        dev = w.device
        self.hi = torch.empty((N, K), dtype=torch.bfloat16, device=dev)
        self.lo = torch.empty_like(self.hi)
        self.hi_t = torch.empty((K, N), dtype=torch.bfloat16, device=dev)
        self.lo_t = torch.empty_like(self.hi_t)
        self.stamp = None

    def fresh(self, w):
        stamp = (w._version, WEIGHT_EPOCH[0], w.data_ptr())
        if stamp != self.stamp:
            N, K = w.shape
            _call("vln_split_bf16", _ptr(w), _ptr(self.hi), _ptr(self.lo), _ptr(self.hi_t), _ptr(self.lo_t), N, K,
                  _stream())
            self.stamp = stamp
        return self


def _split_of(w):
    """The cached split of parameter ``w`` — keyed by the tensor OBJECT (validated through a weak
    reference: ids and device addresses are recycled when an agent is rebuilt)."""
    import weakref
    key = id(w)
    ent = _split_cache.get(key)
    if ent is None or ent[0]() is not w:
        if len(_split_cache) > 256:                         # drop entries whose parameter died
            for k in [k for k, (r, _) in _split_cache.items() if r() is None]:
                del _split_cache[k]
        ent = _split_cache[key] = (weakref.ref(w), _SplitWeight(w))
    return ent[1].fresh(w)


def _tc_matmul(x, hi, lo, N, K, bias=None, out=None):
    """x @ W^T (+ bias) on the tcgen05 kernel, W given by its bf16 hi/lo split [N,K]; added to `out` if given."""
    M = x.shape[0]
    acc = out is not None
    if out is None:
        out = torch.empty((M, N), device=x.device, dtype=torch.float32)
    _call("vln_linear_bf16x3", _ptr(hi), _ptr(lo), N, K, _ptr(x), x.stride(0), M, _ptr(bias), _ptr(out), out.stride(0),
          1 if acc else 0, 0, _stream())
    return out


class _LinearTC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, acc):
        x = _f32c(x)
        sw = _split_of(w)
        N, K = w.shape
        y = _tc_matmul(x, sw.hi, sw.lo, N, K, b.detach() if b is not None else None,
                       acc.detach().clone() if acc is not None else None)
        ctx.save_for_backward(x, w)
        ctx.sw, ctx.has_b, ctx.has_acc = sw, b is not None, acc is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = _f32c(dy)
        N, K = w.shape
        sw = ctx.sw
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            if N % 64 == 0:
                dx = _tc_matmul(dy, sw.hi_t, sw.lo_t, K, N)
            else:
                dx = dy @ w
        if ctx.needs_input_grad[1]:
            dw = wgrad(dy, x)
        if ctx.has_b and ctx.needs_input_grad[2]:
            db = dy.sum(0)
        return dx, dw, db, (dy if ctx.has_acc else None)


_announced = set()


def _announce_library(what, why):
    """Say ONCE per call site when a product falls to a library GEMM instead of the package's own kernels."""
    if what not in _announced:
        _announced.add(what)
        import warnings
        warnings.warn("clvln_b200.ops.%s: library GEMM (%s); the tcgen05 kernels take fp32 2-D operands whose inner sizes "
                      "are multiples of 4" % (what, why), RuntimeWarning, stacklevel=3)


WGRAD_TF32 = [True]       # weight-gradient GEMMs ([N x T*B] x [T*B x K], library calls) on TF32 tensor cores


WGRAD_TC = [__import__("os").environ.get("VLN_WGRAD_TC", "1") != "0"]   # hand-written tcgen05 kind::tf32 kernel (csrc/wgrad.cu)
_wgrad_scratch = {}


def _wgrad_ws(dev):
    """Scratch slab of the split-row merge, one per (device, stream): weight gradients of the decoder and of the encoder
    run on different side streams at the same time."""
    key = (dev.index, torch.cuda.current_stream().cuda_stream)
    ws = _wgrad_scratch.get(key)
    if ws is None:
        ws = _wgrad_scratch[key] = torch.empty(4 << 20, device=dev)
    return ws


def _rows_view(t):
    """[R, C] fp32 with unit column stride and 16-byte aligned rows (row stride free) — what the kernel's TMA needs."""
    if t.dim() != 2 or t.stride(1) != 1 or t.stride(0) % 4 or t.data_ptr() % 16 or t.dtype != torch.float32:
        t = t.contiguous().float()
    return t


def wgrad_tc(dy, x, out=None):
    """dW[M,N] = dY^T X on the tcgen05 kernel; accumulated into `out` (a .grad view) when given."""
    dy, x = _rows_view(dy), _rows_view(x)
    R, M = dy.shape
    N = x.shape[1]
    acc = out is not None
    if out is None:
        out = torch.empty((M, N), device=dy.device)
    assert out.stride(1) == 1 and out.stride(0) % 4 == 0
    scratch = _wgrad_ws(dy.device)
    _call("vln_wgrad_tf32", _ptr(dy), dy.stride(0), _ptr(x), x.stride(0), R, M, N, _ptr(out), out.stride(0), 1 if acc else 0,
          _ptr(scratch), scratch.numel(), _stream())
    return out


def dgrad_tc(dy, w):
    """dX[M,K] = dY[M,N] W[N,K] on the tcgen05 kind::tf32 kernel (dY as the K-major operand)."""
    dy, w = _rows_view(dy), _rows_view(w)
    M, N = dy.shape
    K = w.shape[1]
    out = torch.empty((M, K), device=dy.device)
    _call("vln_dgrad_tf32", _ptr(dy), dy.stride(0), _ptr(w), w.stride(0), M, N, K, _ptr(out), out.stride(0), _stream())
    return out


def seq_outer_sum(a, v, out=None):
    """out[b,l,j] (+)= sum_t a[t,b,l] v[t,b,j] (fp32): a [n,B,L], v [n,B,H] views with unit last stride -> [B,L,H]."""
    n, B, L = a.shape
    H = v.shape[2]
    assert a.stride(2) == 1 and v.stride(2) == 1 and v.shape[:2] == (n, B)
    acc = out is not None
    if out is None:
        out = torch.empty((B, L, H), device=a.device)
    assert out.is_contiguous()
    _call("vln_seq_outer_sum", _ptr(a), a.stride(0), a.stride(1), _ptr(v), v.stride(0), v.stride(1), n, B, L, H, _ptr(out),
          1 if acc else 0, _stream())
    return out


def wgrad(dy, x):
    """dW = dY^T X over all stacked rows (one GEMM per weight and iteration) with TF32 inputs and fp32 accumulation: the
    rounding of the 10-bit mantissas averages out over the T*B-long sums (measured gradient cosine vs the fp32 oracle stays
    >= 0.9999, tests/test_agents_gpu.py).  The hand-written tcgen05 kernel (csrc/wgrad.cu) when the shapes allow it (M, N
    multiples of 4); with WGRAD_TF32 off (the fp32 comparison of the kernel tests) the library fp32 GEMM."""
    if WGRAD_TC[0] and WGRAD_TF32[0] and dy.is_cuda and dy.shape[1] % 4 == 0 and x.shape[1] % 4 == 0:
        return wgrad_tc(dy, x)
    if WGRAD_TF32[0]:
        _announce_library("wgrad", "dY^T X with shape %s x %s" % (tuple(dy.shape), tuple(x.shape)))
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = bool(WGRAD_TF32[0])
    try:
        return dy.t() @ x
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def _tc_matmul_tall(x, hi, lo, N, K, bias=None):
    """x @ W^T (+ bias) for M > 128 rows: 128x128 output blocks, plain stores (csrc/gemm.cu, tall mode)."""
    M = x.shape[0]
    out = torch.empty((M, N), device=x.device, dtype=torch.float32)
    _call("vln_linear_bf16x3_tall", _ptr(hi), _ptr(lo), N, K, _ptr(x), x.stride(0), M, _ptr(bias), _ptr(out),
          out.stride(0), 0, _stream())
    return out


ASYNC_LEAF_GRADS = [False]   # set by TrainStep: leaf (weight) gradients of the encoder on a side stream, into .grad
_leaf_side = {}


def _leaf_grads_async(pairs, keep):
    """pairs: [(parameter, fn)] — fn() computes that parameter's gradient.  Nothing downstream in the backward pass
    needs a leaf gradient, so they are computed on a side stream (under the kernels that carry the chain: the BPTT
    recurrence, the input-gradient GEMM, the embedding backward) and accumulated straight into ``.grad``; the calling
    stream joins when the backward pass ends.  ``keep`` holds the tensors the side stream reads until then."""
    main = torch.cuda.current_stream()
    dev = main.device
    side = _leaf_side.get(dev)
    if side is None:
        side = _leaf_side[dev] = torch.cuda.Stream(device=dev)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        for q, fn in pairs:
            q.grad.add_(fn())

    def join():
        torch.cuda.current_stream().wait_stream(side)
        keep.clear()
    torch.autograd.Variable._execution_engine.queue_callback(join)


def _async_leaf_ok(*params):
    return ASYNC_LEAF_GRADS[0] and all(isinstance(q, torch.nn.Parameter) and q.grad is not None for q in params)


class _LinearTall(torch.autograd.Function):
    """Tall inputs (encoder input projection [B*L, E], batched critic [T*B, H]) on the same tcgen05 bf16x3
    kernel, forward and input gradient; weight gradient through wgrad().  ``dx_tf32``: the input gradient as a
    library TF32 GEMM instead — for the encoder's input projection, whose dx only feeds the embedding gradient
    (same tolerance argument as wgrad(): 1024-long sums, gradient cosine unaffected) and whose K = 4H = 1024,
    N = E = 256 shape keeps the two-stage tcgen05 pipeline latency-bound (47 us vs 12 us per direction)."""

    @staticmethod
    def forward(ctx, x, w, b, dx_tf32):
        x = _f32c(x)
        sw = _split_of(w)
        N, K = w.shape
        ctx.save_for_backward(x, w)
        ctx.sw, ctx.has_b, ctx.dx_tf32, ctx.w_param = sw, b is not None, bool(dx_tf32), w
        return _tc_matmul_tall(x, sw.hi, sw.lo, N, K, b.detach() if b is not None else None)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = _f32c(dy)
        N, K = w.shape
        dx = None
        if ctx.needs_input_grad[0]:
            if ctx.dx_tf32 and WGRAD_TF32[0] and WGRAD_TC[0] and K % 4 == 0:
                dx = dgrad_tc(dy, w.detach())
            else:
                dx = _tc_matmul_tall(dy, ctx.sw.hi_t, ctx.sw.lo_t, K, N)
        dw = None
        if ctx.needs_input_grad[1]:
            if _async_leaf_ok(ctx.w_param):
                _leaf_grads_async([(ctx.w_param, lambda: wgrad(dy, x))], [dy, x])
            else:
                dw = wgrad(dy, x)
        db = dy.sum(0) if ctx.has_b and ctx.needs_input_grad[2] else None
        return dx, dw, db, None


class _LinearLib(torch.autograd.Function):
    """Library GEMM for shapes the tcgen05 kernel does not take (K not a multiple of 64, e.g. the Follower's
    300-wide embeddings): fp32 forward and input gradient, weight gradient through wgrad()."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        ctx.has_b = b is not None
        return F.linear(x, w, b)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dx = dy @ w if ctx.needs_input_grad[0] else None
        dw = wgrad(dy, x) if ctx.needs_input_grad[1] else None
        db = dy.sum(0) if ctx.has_b and ctx.needs_input_grad[2] else None
        return dx, dw, db


def linear(x, w, b=None, acc=None, dx_tf32=False):
    """y = x W^T + b (+ acc) on the tcgen05 bf16x3 kernel (csrc/gemm.cu): split-K skinny launches for at
    most 128 rows, 128x128 output blocks for taller inputs (encoder input projection, batched critic)."""
    assert x.is_cuda, "ops.linear runs on the GPU only (there is no CPU path in this package)"
    ok = (USE_TC_LINEAR[0] and x.dim() == 2 and w.shape[1] % 64 == 0 and w.shape[0] % 4 == 0
          and x.dtype == torch.float32)
    if (not ok and USE_TC_LINEAR[0] and x.dim() == 2 and x.dtype == torch.float32 and w.shape[0] % 4 == 0
            and w.shape[1] % 64 != 0):
        # inner size not a multiple of 64 (the Follower's 300-wide word embeddings): zero columns up to the next multiple;
        # F.pad is differentiable, so the gradients come back sliced
        pad = 64 - w.shape[1] % 64
        return linear(F.pad(x, (0, pad)), F.pad(w, (0, pad)), b, acc, dx_tf32)
    if USE_TC_LINEAR[0] and x.dim() == 2 and x.dtype == torch.float32 and w.shape[1] % 64 == 0:
        # output width the kernels do not take (the critic's single value): zero weight rows up to the next multiple
        # of 4 (skinny) / 64 (tall launches), the extra outputs sliced away; differentiable like the padding above
        N = w.shape[0]
        need = 4 if x.shape[0] <= 128 else 64
        if N % need != 0:
            pad = need - N % need
            y = linear(x, F.pad(w, (0, 0, 0, pad)), None if b is None else F.pad(b, (0, pad)), None, dx_tf32)[:, :N]
            return y if acc is None else y + acc
    if ok and x.shape[0] <= 128:
        return _LinearTC.apply(x, w, b, acc)
    if ok and w.shape[0] % 64 == 0:
        y = _LinearTall.apply(x, w, b, dx_tf32)
        return y if acc is None else y + acc
    _announce_library("linear", "x %s, weight %s" % (tuple(x.shape), tuple(w.shape)))
    y = _LinearLib.apply(x, w, b) if x.dim() == 2 else F.linear(x, w, b)
    return y if acc is None else y + acc


# ---- bit-exact gathers ---------------------------------------------------------------------------
def gather_pano(store, vp, view):
    vp, view = _i32c(vp), _i32c(view)
    B = vp.shape[0]
    out = torch.empty((B, N_VIEWS, F_DIM), device=vp.device, dtype=torch.float32)
    _call("vln_gather_pano", store.handle, _ptr(vp), _ptr(view), _ptr(store.loc4), _ptr(out), B, _stream())
    return out


def gather_cand(store, vp, view, C_slots=NSLOT):
    vp, view = _i32c(vp), _i32c(view)
    B = vp.shape[0]
    out = torch.empty((B, C_slots, F_DIM), device=vp.device, dtype=torch.float32)
    lens = torch.empty((B,), device=vp.device, dtype=torch.int32)
    _call("vln_gather_cand", store.handle, _ptr(vp), _ptr(view), _ptr(store.cand_view), _ptr(store.cand_ang4),
          _ptr(store.n_cand), _ptr(out), _ptr(lens), B, C_slots, _stream())
    return out, lens


def gather_action_feat(store, vp, view, action, ended=None):
    """a_t_prev = cands[b, max(cpu_a_t,0)] (follower.py:141-164, monitor.py:177-191): [B,2176]."""
    vp, view, action = _i32c(vp), _i32c(view), _i32c(action)
    B = vp.shape[0]
    out = torch.empty((B, F_DIM), device=vp.device, dtype=torch.float32)
    _call("vln_gather_action_feat", store.handle, _ptr(vp), _ptr(view), _ptr(action), _ptr(ended), _ptr(store.cand_view),
          _ptr(store.cand_ang4), _ptr(store.n_cand), _ptr(out), B, _stream())
    return out


def pose_feature(store, view):
    """make_angle_feat(heading, elevation) of the agent's pose (envdrop.py:76-78): [B,128]."""
    return store.pose128[view.long()]


# ---- fused gather + panorama attention ---------------------------------------------------------
def pano_attn_raw(store, vp, view, vec, attn, mode, drop_p=0.0, rng=None, call_off=0, split=1, fwd_out=None):
    B = vp.shape[0]
    out = torch.empty((B, F_DIM), device=vec.device, dtype=torch.float32)
    _call("vln_pano_attn", store.handle, _ptr(vp), _ptr(view), _ptr(store.loc4), _ptr(vec), _ptr(attn),
          _ptr(fwd_out), _ptr(out), B, mode, float(drop_p), rng.ptr if rng is not None else None, call_off, split,
          _stream())
    return out


class _PanoAttn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, store, vp, view, drop_p, rng, call_off, split):
        q = _f32c(q)
        attn = torch.empty((q.shape[0], N_VIEWS), device=q.device, dtype=torch.float32)
        out = pano_attn_raw(store, vp, view, q, attn, 0, drop_p, rng, call_off, split)
        ctx.store, ctx.cfg = store, (drop_p, rng, call_off, split)
        ctx.save_for_backward(vp, view, attn, out)
        ctx.mark_non_differentiable(attn)
        return out, attn

    @staticmethod
    def backward(ctx, d_out, _d_attn):
        vp, view, attn, out = ctx.saved_tensors
        drop_p, rng, call_off, split = ctx.cfg
        dq = pano_attn_raw(ctx.store, vp, view, _f32c(d_out), attn, 1, drop_p, rng, call_off, split, out)
        return dq, None, None, None, None, None, None, None


def pano_attn(store, vp, view, q, drop_p=0.0, rng=None, call_off=0, split=1):
    """(weighted [B,2176], attn [B,36]) = softmax_v(x~_v . q) over the episode's panorama."""
    return _PanoAttn.apply(q, store, _i32c(vp), _i32c(view), drop_p, rng, call_off, split)


# ---- candidate logits -----------------------------------------------------------------------------
class _CandLogits(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tgt, bias, store, vp, view, drop_p, rng, call_off):
        tgt = _f32c(tgt)
        bias = _f32c(bias) if bias is not None else None
        B = tgt.shape[0]
        logits = torch.empty((B, NSLOT), device=tgt.device, dtype=torch.float32)
        _call("vln_cand_logits_fwd", store.handle, _ptr(vp), _ptr(view), _ptr(store.cand_view),
              _ptr(store.cand_ang4), _ptr(store.n_cand), _ptr(tgt), _ptr(bias), _ptr(logits), B, float(drop_p),
              rng.ptr if rng is not None else None, call_off, _stream())
        ctx.store, ctx.cfg, ctx.has_bias = store, (drop_p, rng, call_off), bias is not None
        ctx.save_for_backward(vp, view)
        return logits

    @staticmethod
    def backward(ctx, d_logits):
        vp, view = ctx.saved_tensors
        drop_p, rng, call_off = ctx.cfg
        store = ctx.store
        d_logits = _f32c(d_logits)
        B = d_logits.shape[0]
        d_tgt = torch.empty((B, F_DIM), device=d_logits.device, dtype=torch.float32)
        d_bias = torch.empty((B,), device=d_logits.device, dtype=torch.float32) if ctx.has_bias else None
        _call("vln_cand_logits_bwd", store.handle, _ptr(vp), _ptr(view), _ptr(store.cand_view),
              _ptr(store.cand_ang4), _ptr(store.n_cand), _ptr(d_logits), _ptr(d_tgt), _ptr(d_bias), B,
              float(drop_p), rng.ptr if rng is not None else None, call_off, _stream())
        return d_tgt, d_bias, None, None, None, None, None, None


def cand_logits(store, vp, view, tgt, bias=None, drop_p=0.0, rng=None, call_off=0):
    """[B,16] masked candidate logits straight from the table (no [B,C,2176] tensor)."""
    return _CandLogits.apply(tgt, bias, store, _i32c(vp), _i32c(view), drop_p, rng, call_off)


# ---- soft-dot attention over a dense context (instruction ctx, or a materialised feature tensor) ----
class _CtxAttn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, context, tgt, lengths):
        context, tgt = _f32c(context), _f32c(tgt)
        B, L, H = context.shape
        attn = torch.empty((B, L), device=tgt.device, dtype=torch.float32)
        weighted = torch.empty((B, H), device=tgt.device, dtype=torch.float32)
        _call("vln_ctx_attn_fwd", _ptr(context), _ptr(tgt), _ptr(lengths), _ptr(attn), _ptr(weighted), B, L, H,
              _stream())
        ctx.save_for_backward(context, tgt, lengths, attn)
        return weighted, attn

    @staticmethod
    def backward(ctx, d_weighted, d_attn):
        context, tgt, lengths, attn = ctx.saved_tensors
        B, L, H = context.shape
        d_tgt = torch.empty_like(tgt)
        d_context = torch.zeros_like(context) if ctx.needs_input_grad[0] else None
        d_weighted = _f32c(d_weighted) if d_weighted is not None else torch.zeros_like(tgt)
        d_attn = _f32c(d_attn) if d_attn is not None else None
        _call("vln_ctx_attn_bwd", _ptr(context), _ptr(tgt), _ptr(lengths), _ptr(attn), _ptr(d_weighted),
              _ptr(d_attn), _ptr(d_tgt), _ptr(d_context), B, L, H, _stream())
        return d_context, d_tgt, None


def ctx_attn(context, tgt, lengths):
    """(weighted [B,H], attn [B,L]): softmax over the first lengths[b] rows of context[b]."""
    return _CtxAttn.apply(context, tgt, _i32c(lengths))


def mask_to_lengths(mask, L):
    """The reference passes boolean pad masks (True = masked, always a suffix: base.py:128,
    misc.py:481-486); the kernels take the number of valid rows."""
    if mask is None:
        return None
    return (L - mask.sum(1)).to(torch.int32)


# ---- LSTM pointwise ----------------------------------------------------------------------------------
class _LstmPointwise(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gates, c0):
        gates, c0 = _f32c(gates), _f32c(c0)
        B, H = c0.shape
        h1, c1, acts = torch.empty_like(c0), torch.empty_like(c0), torch.empty_like(gates)
        _call("vln_lstm_pointwise_fwd", _ptr(gates), _ptr(c0), _ptr(h1), _ptr(c1), _ptr(acts), B, H, _stream())
        ctx.save_for_backward(acts, c0, c1)
        return h1, c1

    @staticmethod
    def backward(ctx, d_h1, d_c1):
        acts, c0, c1 = ctx.saved_tensors
        B, H = c0.shape
        d_gates, d_c0 = torch.empty_like(acts), torch.empty_like(c0)
        d_h1 = _f32c(d_h1) if d_h1 is not None else None
        d_c1 = _f32c(d_c1) if d_c1 is not None else None
        _call("vln_lstm_pointwise_bwd", _ptr(acts), _ptr(c0), _ptr(c1), _ptr(d_h1), _ptr(d_c1), _ptr(d_gates),
              _ptr(d_c0), B, H, _stream())
        return d_gates, d_c0


def lstm_pointwise(gates, c0):
    return _LstmPointwise.apply(gates, c0)


def lstm_cell(x, h, c, w_ih, w_hh, b_ih, b_hh):
    """nn.LSTMCell (policy.py:53,159,238): gate GEMMs + fused pointwise kernel."""
    gates = linear(h, w_hh, None, acc=linear(x, w_ih, b_ih + b_hh))
    return lstm_pointwise(gates, c)


def lstm_sequence(xproj, lengths, w_hh, reverse):
    """One direction of a packed-sequence LSTM (units.py:58-71) with length masking instead of
    packing: rows stop at their own length (reverse rows start there), outputs past the length are
    zero, final (h, c) are the states at each row's last valid step.  ``xproj`` [B,L,4H] already
    holds x W_ih^T + b_ih + b_hh."""
    B, L, H4 = xproj.shape
    H = H4 // 4
    h = xproj.new_zeros(B, H)
    c = xproj.new_zeros(B, H)
    outs = [None] * L
    order = range(L - 1, -1, -1) if reverse else range(L)
    live_all = (torch.arange(L, device=xproj.device).unsqueeze(0) < lengths.unsqueeze(1))    # [B,L]
    for t in order:
        live = live_all[:, t:t + 1]
        gates = xproj[:, t] + linear(h, w_hh)
        h_new, c_new = lstm_pointwise(gates, c)
        h = torch.where(live, h_new, h)
        c = torch.where(live, c_new, c)
        outs[t] = h_new * live
    return torch.stack(outs, 1), h, c


def _ptr_array(tensors):
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


class _LstmLayer(torch.autograd.Function):
    """One (bi)directional LSTM layer over padded sequences: the persistent cluster kernel forward,
    its BPTT twin backward; the weight gradient of W_hh is ONE GEMM over all timesteps."""

    @staticmethod
    def forward(ctx, lengths, n_dir, *tensors):
        xproj = [_f32c(t) for t in tensors[:n_dir]]
        w_hh = [_f32c(t) for t in tensors[n_dir:2 * n_dir]]
        B, L, H4 = xproj[0].shape
        H = H4 // 4
        dev = xproj[0].device
        out = torch.zeros((B, L, n_dir * H), device=dev)
        h_last, c_last = torch.empty((B, n_dir * H), device=dev), torch.empty((B, n_dir * H), device=dev)
        acts = [torch.empty((B, L, H4), device=dev) for _ in range(n_dir)]
        cs = [torch.empty((B, L, H), device=dev) for _ in range(n_dir)]
        _call("vln_lstm_seq_fwd", _ptr_array(xproj), _ptr_array(w_hh), _ptr(lengths), _ptr(out), _ptr_array(acts),
              _ptr_array(cs), _ptr(h_last), _ptr(c_last), B, L, H, n_dir, _stream())
        ctx.n_dir = n_dir
        ctx.w_params = tuple(tensors[n_dir:2 * n_dir])
        ctx.save_for_backward(lengths, out, *w_hh, *acts, *cs)
        return out, h_last, c_last

    @staticmethod
    def backward(ctx, d_out, d_h, d_c):
        n = ctx.n_dir
        sv = ctx.saved_tensors
        lengths, out = sv[0], sv[1]
        w_hh, acts, cs = sv[2:2 + n], sv[2 + n:2 + 2 * n], sv[2 + 2 * n:2 + 3 * n]
        B, L, H4 = acts[0].shape
        H = H4 // 4
        d_x = [torch.zeros_like(a) for a in acts]
        d_out = _f32c(d_out) if d_out is not None else None
        d_h = _f32c(d_h) if d_h is not None else None
        d_c = _f32c(d_c) if d_c is not None else None
        _call("vln_lstm_seq_bwd", _ptr_array(w_hh), _ptr(lengths), _ptr_array(acts), _ptr_array(cs), _ptr(d_out),
              _ptr(d_h), _ptr(d_c), _ptr_array(d_x), B, L, H, n, _stream())
        def dw_of(k):
            hk = out[:, :, k * H:(k + 1) * H]
            hprev = torch.zeros_like(hk)
            if k == 0:
                hprev[:, 1:] = hk[:, :-1]                       # h_{t-1}; zero initial state
            else:
                hprev[:, :-1] = hk[:, 1:]                       # reversed direction: the previous step is t+1
            return wgrad(d_x[k].reshape(B * L, H4), hprev.reshape(B * L, H))
        if _async_leaf_ok(*ctx.w_params):
            _leaf_grads_async([(ctx.w_params[k], (lambda k=k: dw_of(k))) for k in range(n)], [out, d_x])
            d_w = [None] * n
        else:
            d_w = [dw_of(k) for k in range(n)]
        return (None, None, *d_x, *d_w)


def lstm_layer(xproj, w_hh, lengths):
    """xproj / w_hh: lists with one entry per direction.  -> (out [B,L,n_dir*H], h_last, c_last)."""
    return _LstmLayer.apply(_i32c(lengths), len(xproj), *xproj, *w_hh)


LSTM_KERNEL_H = (128, 256, 512)


# ---- action head ---------------------------------------------------------------------------------------
FEEDBACK = {"teacher": 0, "argmax": 1, "sample": 2}


class _Policy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, feedback, rng, call_off):
        logits = _f32c(logits)
        B = logits.shape[0]
        dev = logits.device
        ce = torch.empty((B,), device=dev, dtype=torch.float32)
        logp, ent = torch.empty_like(ce), torch.empty_like(ce)
        action = torch.empty((B,), device=dev, dtype=torch.int32)
        probs = torch.empty((B, NSLOT), device=dev, dtype=torch.float32)
        _call("vln_policy_fwd", _ptr(logits), _ptr(target), feedback, rng.ptr if rng is not None else None,
              call_off, _ptr(ce), _ptr(action), _ptr(logp), _ptr(ent), _ptr(probs), B, _stream())
        ctx.save_for_backward(probs, target, action, ent)
        ctx.mark_non_differentiable(action)
        return ce, logp, ent, action

    @staticmethod
    def backward(ctx, g_ce, g_logp, g_ent, _g_action):
        probs, target, action, ent = ctx.saved_tensors
        B = probs.shape[0]
        d = torch.empty_like(probs)
        g = [_f32c(x) if x is not None else None for x in (g_ce, g_logp, g_ent)]
        if target is None:
            g[0] = None
        _call("vln_policy_bwd", _ptr(probs), _ptr(target), _ptr(action), _ptr(ent), _ptr(g[0]), _ptr(g[1]),
              _ptr(g[2]), _ptr(d), B, _stream())
        return d, None, None, None, None


def policy_head(logits, target, feedback, rng=None, call_off=0):
    """logits [B,16] (-inf = masked) -> (ce [B], logp [B], entropy [B], action int32 [B])."""
    if logits.shape[1] != NSLOT:
        pad = logits.new_full((logits.shape[0], NSLOT - logits.shape[1]), float("-inf"))
        logits = torch.cat((logits, pad), 1)
    fb = FEEDBACK[feedback] if isinstance(feedback, str) else int(feedback)
    return _Policy.apply(logits, _i32c(target) if target is not None else None, fb, rng, call_off)


# ---- dropout ----------------------------------------------------------------------------------------------
class _Dropout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p, rng, call_off):
        x = _f32c(x)
        y = torch.empty_like(x)
        _call("vln_dropout", _ptr(x), _ptr(y), x.numel(), float(p), rng.ptr, call_off, _stream())
        ctx.cfg = (p, rng, call_off)
        return y

    @staticmethod
    def backward(ctx, g):
        p, rng, call_off = ctx.cfg
        g = _f32c(g)
        y = torch.empty_like(g)
        _call("vln_dropout", _ptr(g), _ptr(y), g.numel(), float(p), rng.ptr, call_off, _stream())
        return y, None, None, None


def dropout(x, p, rng, tag="drop"):
    """nn.Dropout on the library's Philox stream; identity when p == 0 (eval)."""
    if p <= 0.0:
        return x
    return _Dropout.apply(x, p, rng, rng.next(tag, x.shape, p))


class _EmbedDrop(torch.autograd.Function):
    """nn.Embedding(padding_idx) -> nn.Dropout (units.py:48-52) as one kernel each way."""

    @staticmethod
    def forward(ctx, tokens, weight, padding_idx, p, rng, call_off):
        tokens = tokens.contiguous()
        assert tokens.is_cuda and tokens.dtype == torch.int64
        w = _f32c(weight)
        V, E = w.shape
        y = torch.empty(tokens.shape + (E,), device=w.device, dtype=torch.float32)
        _call("vln_embed_drop_fwd", _ptr(tokens), _ptr(w), _ptr(y), tokens.numel(), E, V, float(p), rng.ptr, call_off,
              _stream())
        ctx.save_for_backward(tokens)
        ctx.cfg = (V, E, -1 if padding_idx is None else int(padding_idx), float(p), rng, call_off)
        return y

    @staticmethod
    def backward(ctx, g):
        tokens, = ctx.saved_tensors
        V, E, pad, p, rng, call_off = ctx.cfg
        g = _f32c(g)
        d_w = torch.empty((V, E), device=g.device, dtype=torch.float32)
        _call("vln_embed_drop_bwd", _ptr(tokens), _ptr(g), _ptr(d_w), tokens.numel(), E, V, pad, p, rng.ptr, call_off,
              _stream())
        return None, d_w, None, None, None, None


def embed_dropout(tokens, weight, padding_idx, p, rng, tag="enc_embed"):
    """drop(embedding(tokens)); p == 0 (eval) is the plain lookup through the same kernel."""
    off = rng.next(tag, tuple(tokens.shape) + (weight.shape[1],), p) if p > 0.0 else 0
    return _EmbedDrop.apply(tokens, weight, padding_idx, p, rng, off)


def dropout_mask(shape, p, rng, call_off, device=None):
    """The keep-mask (uint8) the kernels use for a tensor of this shape under (rng, call_off)."""
    m = torch.empty(shape, device=device or rng.state.device, dtype=torch.uint8)
    _call("vln_dropout_mask", _ptr(m), m.numel(), float(p), rng.ptr, call_off, _stream())
    return m


# ---- environment ---------------------------------------------------------------------------------------------
def env_observe(store, vp, ended, goal):
    B = vp.shape[0]
    teacher = torch.empty((B,), device=vp.device, dtype=torch.int32)
    dist = torch.empty((B,), device=vp.device, dtype=torch.float32)
    _call("vln_env_observe", _ptr(vp), _ptr(ended), _ptr(goal), _ptr(store.cand_vp), _ptr(store.n_cand),
          _ptr(store.next_hop), _ptr(store.dist), _ptr(store.sq_off), _ptr(store.vp_local), _ptr(teacher),
          _ptr(dist), B, _stream())
    return teacher, dist


def env_step(store, vp, view, ended, dist, goal, action, out=None, n_active=None):
    """(vp', view', ended', dist', teacher', reward, mask) for one transition; inputs untouched.
    ``out`` may supply preallocated (vp', view', ended', dist') rows of a trajectory buffer."""
    B = vp.shape[0]
    dev = vp.device
    if out is None:
        out = (torch.empty_like(vp), torch.empty_like(view), torch.empty_like(ended), torch.empty_like(dist))
    vp2, view2, ended2, dist2 = out
    teacher = torch.empty((B,), device=dev, dtype=torch.int32)
    reward = torch.empty((B,), device=dev, dtype=torch.float32)
    mask = torch.empty((B,), device=dev, dtype=torch.float32)
    _call("vln_env_step", _ptr(vp), _ptr(view), _ptr(ended), _ptr(dist), _ptr(goal), _ptr(_i32c(action)),
          _ptr(store.cand_vp), _ptr(store.cand_view), _ptr(store.n_cand), _ptr(store.next_hop), _ptr(store.dist),
          _ptr(store.sq_off), _ptr(store.vp_local), _ptr(vp2), _ptr(view2), _ptr(ended2), _ptr(dist2),
          _ptr(teacher), _ptr(reward), _ptr(mask), _ptr(n_active), B, _stream())
    return vp2, view2, ended2, dist2, teacher, reward, mask


# ---- A2C ---------------------------------------------------------------------------------------------------------
class _A2C(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logp, entropy, value, reward, mask, last_value, ended, gamma, ent_coef):
        T, B = reward.shape
        dev = reward.device
        logp, value = _f32c(logp), _f32c(value)
        entropy = _f32c(entropy) if entropy is not None else None
        loss_b = torch.empty((B,), device=dev, dtype=torch.float32)
        ret = torch.empty((T, B), device=dev, dtype=torch.float32)
        stats = torch.zeros((2,), device=dev, dtype=torch.float32)
        _call("vln_a2c_fwd", _ptr(reward), _ptr(mask), _ptr(logp), _ptr(entropy), _ptr(value), _ptr(last_value),
              _ptr(ended), float(gamma), float(ent_coef), _ptr(loss_b), _ptr(ret), _ptr(stats),
              C.c_void_p(stats.data_ptr() + 4), T, B, _stream())
        ctx.save_for_backward(mask, value, ret)
        ctx.ent_coef = float(ent_coef)
        ctx.has_ent = entropy is not None
        ctx.mark_non_differentiable(stats)
        return loss_b, stats

    @staticmethod
    def backward(ctx, g_b, _g_stats):
        mask, value, ret = ctx.saved_tensors
        T, B = mask.shape
        d_logp, d_value = torch.empty_like(value), torch.empty_like(value)
        d_ent = torch.empty_like(value) if ctx.has_ent else None
        _call("vln_a2c_bwd", _ptr(_f32c(g_b)), _ptr(mask), _ptr(value), _ptr(ret), ctx.ent_coef, _ptr(d_logp),
              _ptr(d_value), _ptr(d_ent), T, B, _stream())
        return d_logp, d_ent, d_value, None, None, None, None, None, None


def a2c_loss(logp, entropy, value, reward, mask, last_value, ended, gamma, ent_coef=0.01):
    """Per-episode A2C loss [B] and stats [2] = (sum of masks, sum of masked critic errors^2)
    (envdrop.py:240-264); all step tensors are time-major [T,B]."""
    return _A2C.apply(logp, entropy, value, _f32c(reward), _f32c(mask), _f32c(last_value), ended.contiguous(),
                      gamma, ent_coef if entropy is not None else 0.0)
