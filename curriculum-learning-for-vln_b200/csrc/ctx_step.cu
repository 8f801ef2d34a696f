// Text-attention stage of a fused EnvDrop decoder step as ONE launch each way
// (EnvDropDecoder.forward policy.py:238-241: nn.LSTMCell pointwise half, nn.Dropout on h_1, then
// SoftDotAttention over the instruction units.py:107-118 up to the weighted context).
//
// Replaces three launches of the step chain (lstm_pw_drop_fwd -> tq GEMM -> ctx_attn, and in backward
// ctx_attn -> tin^T GEMM -> lstm_pw_drop_bwd) by one, using the identity
//     logit_l = ctx_l . (W_in h) = (ctx W_in)_l . h = CW_l . h ,        CW = ctx W_in  [B, L, H]
// CW is one tall GEMM per rollout (the instruction context is constant over the decoder steps), so the
// per-step product W_in h — a grid-wide GEMM launch on the latency chain — disappears; its gradient returns as
//     d_h += sum_l dlogit_l CW_l        (this kernel)        dCW = sum_t dlogit_t (x) h_t   (one bmm per rollout).
//
// One CTA per episode, 512 threads = the 512 hidden units.  Both [len, 512] fp32 tiles of the episode are
// constant over the rollout and are requested BEFORE the programmatic-dependency wait:
//   * the tile that is DOTTED with a vector (CW forward, ctx backward) lands in shared memory by one bulk copy
//     (one warp per row, conflict-free 16-byte reads);
//   * the tile that is WEIGHTED-SUMMED over rows (ctx forward, CW backward) lives in REGISTERS: thread u holds
//     column u of all <= 80 rows (coalesced 2 KB row reads), so out_u = sum_l w_l x_lu is 80 register FMAs.
// Thread u also owns hidden unit u of the LSTM pointwise half: forward it turns the gate pre-activations into
// (h_1, c_1), saves the activations and writes drop(h_1) into the [weighted | h] operand row of linear_out;
// backward it finishes d(drop(h_1)) and writes the four gate gradients + d_c0.
// Roofline class: L2 — 2 x len x 512 x 4 B per episode-step each way (+ 20 KB of gates / activations).
#include "common.cuh"

namespace {

constexpr int kH = 512;
constexpr int kThreads = kH;
constexpr int kWarps = kThreads / 32;
constexpr int kLmax = 80;                               // rows held in registers (instructions are capped at 80 tokens)

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

struct Smem {
  float tile[kLmax * kH];                               // 160 KB: the dotted tile
  float vec[kH];                                        // drop(h_1) (forward) / d_weighted (backward)
  float sc[kLmax + 16];                                 // logits -> attention (forward) / r_l -> dlogit (backward)
  float av[kLmax + 16];                                 // saved attention (backward)
  uint64_t bar;
};

// keep-scale (0 or 1/(1-p)) of element (b, u) of a dense [B, H] tensor under stream (rng, call_off)
__device__ __forceinline__ float keep_of(float p, const uint64_t* rng, uint64_t call_off, int b, int u) {
  if (!(p > 0.f)) return 1.f;
  const Philox8 r = philox8(rng[0], rng[1] + call_off, ((uint64_t)b * kH + (uint64_t)u) >> 3);
  return philox_keep(r, u & 7, drop_threshold(p)) ? 1.0f / (1.0f - p) : 0.f;
}

// rows [0, len) of `tile` dotted with vec -> sc[l]; one warp per row
__device__ __forceinline__ void row_dots(const float* tile, const float* vec, float* sc, int len, int warp, int lane) {
  const float4* v4 = reinterpret_cast<const float4*>(vec);
  for (int l = warp; l < len; l += kWarps) {
    const float4* row = reinterpret_cast<const float4*>(tile + (size_t)l * kH);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int j = 0; j < kH / 128; ++j) {                 // (the column tile of this thread keeps ~80 registers busy:
      const float4 x = row[j * 32 + lane], q = v4[j * 32 + lane];      //  the vector is re-read from shared memory)
      a0 = fmaf(x.x, q.x, a0); a1 = fmaf(x.y, q.y, a1); a2 = fmaf(x.z, q.z, a2); a3 = fmaf(x.w, q.w, a3);
    }
    const float s = warp_sum((a0 + a1) + (a2 + a3));
    if (lane == 0) sc[l] = s;
  }
}

// sum_l sc[l] * col[l] with four independent chains (sc[l] = 0 for l >= len)
__device__ __forceinline__ float col_weighted(const float* sc, const float (&col)[kLmax]) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
  for (int l = 0; l < kLmax; l += 4) {
    const float4 w = *reinterpret_cast<const float4*>(sc + l);
    a0 = fmaf(w.x, col[l], a0); a1 = fmaf(w.y, col[l + 1], a1);
    a2 = fmaf(w.z, col[l + 2], a2); a3 = fmaf(w.w, col[l + 3], a3);
  }
  return (a0 + a1) + (a2 + a3);
}

struct FwdArgs {
  const float* gates; const float* c0; float* h1; float* c1; float* acts; float* wh; int ld_wh;
  const float* ctx; const float* cw; const int32_t* lengths; float* attn;
  int L; float p; const uint64_t* rng; uint64_t call_off; int ready;
};

__global__ void __launch_bounds__(kThreads, 1) ctx_step_fwd_kernel(const __grid_constant__ FwdArgs a,
                                                                   const __grid_constant__ ChainLink link) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int u = threadIdx.x, lane = u & 31, warp = u >> 5, b = blockIdx.x, L = a.L;
  CHAIN_BEGIN(a.rng, 3);
  pdl_trigger();
  // ready: ctx / CW / lengths were complete before the preceding kernel started (every decoder step but the first),
  // so both tiles are requested before the dependency wait and land while the gates GEMM is still finishing
  if (!a.ready) chain_wait_cta(link);
  const int len = max(0, min(a.lengths[b], min(L, kLmax)));
  if (u == 0) {
    mbar_init(&sm.bar, 1);
    fence_mbar_init();
    if (len > 0) {
      mbar_expect_tx(&sm.bar, (uint32_t)len * kH * 4u);
      bulk_g2s(sm.tile, a.cw + (size_t)b * L * kH, (uint32_t)len * kH * 4u, &sm.bar);
    }
  }
  float col[kLmax];                                      // column u of the episode's context rows
  {
    const float* cb = a.ctx + (size_t)b * L * kH + u;
#pragma unroll
    for (int l = 0; l < kLmax; ++l) col[l] = l < len ? __ldg(cb + (size_t)l * kH) : 0.f;
  }
  if (u < kLmax + 16) sm.sc[u] = 0.f;
  const float keep = keep_of(a.p, a.rng, a.call_off, b, u);     // (the Philox base moves only between iterations)
  if (a.ready) chain_wait_cta(link);
  CHAIN_MARK(2);
  // ---- nn.LSTMCell pointwise half for hidden unit u (gate order i, f, g, o) + dropout of h_1 ----
  {
    const float* gr = a.gates + (size_t)b * 4 * kH + u;
    const float gi = sigmoidf_(gr[0]), gf = sigmoidf_(gr[kH]), gg = tanhf(gr[2 * kH]), go = sigmoidf_(gr[3 * kH]);
    const size_t i = (size_t)b * kH + u;
    const float c = gf * a.c0[i] + gi * gg;
    const float h = go * tanhf(c);
    a.c1[i] = c;
    a.h1[i] = h;
    float* ar = a.acts + (size_t)b * 4 * kH + u;
    ar[0] = gi; ar[kH] = gf; ar[2 * kH] = gg; ar[3 * kH] = go;
    const float hd = h * keep;
    a.wh[(size_t)b * a.ld_wh + kH + u] = hd;
    sm.vec[u] = hd;
  }
  __syncthreads();                                       // barrier initialised, vec complete
  if (len > 0) mbar_wait(&sm.bar, 0);
  row_dots(sm.tile, sm.vec, sm.sc, len, warp, lane);     // logit_l = CW_l . drop(h_1)
  __syncthreads();
  if (warp == 0) {                                       // masked softmax over the first len rows (units.py:112-115)
    float x[3], e[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) x[k] = (lane + 32 * k < len) ? sm.sc[lane + 32 * k] : -INFINITY;
    const float m = warp_max(fmaxf(x[0], fmaxf(x[1], x[2])));
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      e[k] = (lane + 32 * k < len) ? expf(x[k] - m) : 0.f;
      s += e[k];
    }
    const float inv = 1.0f / warp_sum(s);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int l = lane + 32 * k;
      const float pr = l < len ? e[k] * inv : 0.f;
      if (l < kLmax) sm.sc[l] = pr;
      if (l < L) a.attn[(size_t)b * L + l] = pr;
    }
  }
  __syncthreads();
  a.wh[(size_t)b * a.ld_wh + u] = col_weighted(sm.sc, col);      // weighted context, column u
  chain_signal_cta(link);
  CHAIN_MARK(3);
}

struct BwdArgs {
  const float* ctx; const float* cw; const int32_t* lengths; const float* attn; const float* dwh; int ld_dwh;
  float* dlogit; const float* acts; const float* c0; const float* c1; const float* d_h1_extra; const float* d_c1;
  float* d_gates; float* d_c0; int L; float p; const uint64_t* rng; uint64_t call_off;
};

__global__ void __launch_bounds__(kThreads, 1) ctx_step_bwd_kernel(const __grid_constant__ BwdArgs a,
                                                                   const __grid_constant__ ChainLink link) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int u = threadIdx.x, lane = u & 31, warp = u >> 5, b = blockIdx.x, L = a.L;
  CHAIN_BEGIN(a.rng, 13);
  pdl_trigger();
  // everything read before the wait dates from the forward pass (tiles, lengths, saved attention / activations / cells)
  const int len = max(0, min(a.lengths[b], min(L, kLmax)));
  if (u == 0) {
    mbar_init(&sm.bar, 1);
    fence_mbar_init();
    if (len > 0) {
      mbar_expect_tx(&sm.bar, (uint32_t)len * kH * 4u);
      bulk_g2s(sm.tile, a.ctx + (size_t)b * L * kH, (uint32_t)len * kH * 4u, &sm.bar);
    }
  }
  float col[kLmax];                                      // column u of CW
  {
    const float* cb = a.cw + (size_t)b * L * kH + u;
#pragma unroll
    for (int l = 0; l < kLmax; ++l) col[l] = l < len ? __ldg(cb + (size_t)l * kH) : 0.f;
  }
  if (u < kLmax + 16) {
    sm.sc[u] = 0.f;
    sm.av[u] = u < len ? a.attn[(size_t)b * L + u] : 0.f;
  }
  const size_t i = (size_t)b * kH + u;
  const float* ar = a.acts + (size_t)b * 4 * kH + u;
  const float gi = ar[0], gf = ar[kH], gg = ar[2 * kH], go = ar[3 * kH];
  const float cp = a.c0[i], cn = a.c1[i];
  const float keep = keep_of(a.p, a.rng, a.call_off, b, u);
  const float tc = tanhf(cn);
  chain_wait_cta(link);
  CHAIN_MARK(2);
  sm.vec[u] = a.dwh[(size_t)b * a.ld_dwh + u];           // d_weighted (from the linear_out input-gradient GEMM)
  const float dhd_gemm = a.dwh[(size_t)b * a.ld_dwh + kH + u];
  __syncthreads();
  if (len > 0) mbar_wait(&sm.bar, 0);
  row_dots(sm.tile, sm.vec, sm.sc, len, warp, lane);     // r_l = ctx_l . d_weighted
  __syncthreads();
  if (warp == 0) {                                       // dlogit = attn * (r - attn . r)
    float x[3], pr[3];
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int l = lane + 32 * k;
      x[k] = l < len ? sm.sc[l] : 0.f;
      pr[k] = l < len ? sm.av[l] : 0.f;
      dot += pr[k] * x[k];
    }
    dot = warp_sum(dot);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int l = lane + 32 * k;
      const float d = l < len ? pr[k] * (x[k] - dot) : 0.f;
      if (l < kLmax) sm.sc[l] = d;
      if (l < L) a.dlogit[(size_t)b * L + l] = d;
    }
  }
  __syncthreads();
  // d(drop(h_1))_u = (W_out^T dpre)_u [GEMM] + sum_l dlogit_l CW_lu ; through the dropout, + the critic's gradient on h_1
  float dh = (dhd_gemm + col_weighted(sm.sc, col)) * keep;
  if (a.d_h1_extra) dh += a.d_h1_extra[i];
  // ---- LSTMCell pointwise backward ----
  const float dc = (a.d_c1 ? a.d_c1[i] : 0.f) + dh * go * (1.f - tc * tc);
  float* dg = a.d_gates + (size_t)b * 4 * kH + u;
  dg[0] = dc * gg * gi * (1.f - gi);
  dg[kH] = dc * cp * gf * (1.f - gf);
  dg[2 * kH] = dc * gi * (1.f - gg * gg);
  dg[3 * kH] = dh * tc * go * (1.f - go);
  a.d_c0[i] = dc * gf;
  chain_signal_cta(link);
  CHAIN_MARK(3);
}

int configure() {
  static bool done = false;
  if (!done) {
    VLN_CHECK_CUDA(cudaFuncSetAttribute(ctx_step_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
    VLN_CHECK_CUDA(cudaFuncSetAttribute(ctx_step_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
    done = true;
  }
  return 0;
}

}  // namespace

extern "C" int vln_envdrop_ctx_step_fwd(const float* gates, const float* c0, float* h1, float* c1, float* acts, float* wh,
                                        int ld_wh, const float* ctx, const float* cw, const int32_t* lengths, float* attn,
                                        int B, int L, int H, float p, const uint64_t* rng, uint64_t call_off, int tiles_ready,
                                        void* stream) {
  VLN_REQUIRE(gates && c0 && h1 && c1 && acts && wh && ctx && cw && lengths && attn && B > 0, "bad arguments");
  VLN_REQUIRE(H == kH, "hidden size must be 512");
  VLN_REQUIRE(L > 0 && L <= kLmax, "L must be in 1..80");
  VLN_REQUIRE(ld_wh >= 2 * kH, "wh rows hold [weighted | drop(h)]: stride >= 1024");
  VLN_REQUIRE(p >= 0.f && p < 1.f && (p == 0.f || rng), "dropout needs 0 <= p < 1 and an rng state");
  VLN_REQUIRE((((uintptr_t)ctx | (uintptr_t)cw) & 15) == 0, "ctx / cw must be 16-byte aligned");
  if (int rc = configure()) return rc;
  FwdArgs a{gates, c0, h1, c1, acts, wh, ld_wh, ctx, cw, lengths, attn, L, p, rng, call_off, tiles_ready ? 1 : 0};
  const ChainLink link = vln_chain_link((cudaStream_t)stream, (unsigned int)B);
  VLN_CHECK_CUDA(vln_launch_linked(ctx_step_fwd_kernel, dim3(B), dim3(kThreads), sizeof(Smem), (cudaStream_t)stream, a, link));
  return 0;
}

extern "C" int vln_envdrop_ctx_step_bwd(const float* ctx, const float* cw, const int32_t* lengths, const float* attn,
                                        const float* dwh, int ld_dwh, float* dlogit_out, const float* acts, const float* c0,
                                        const float* c1, const float* d_h1_extra, const float* d_c1, float* d_gates,
                                        float* d_c0, int B, int L, int H, float p, const uint64_t* rng, uint64_t call_off,
                                        void* stream) {
  VLN_REQUIRE(ctx && cw && lengths && attn && dwh && dlogit_out && acts && c0 && c1 && d_gates && d_c0 && B > 0,
              "bad arguments");
  VLN_REQUIRE(H == kH, "hidden size must be 512");
  VLN_REQUIRE(L > 0 && L <= kLmax, "L must be in 1..80");
  VLN_REQUIRE(ld_dwh >= 2 * kH, "dwh rows hold [d_weighted | d drop(h)]: stride >= 1024");
  VLN_REQUIRE(p >= 0.f && p < 1.f && (p == 0.f || rng), "dropout needs 0 <= p < 1 and an rng state");
  VLN_REQUIRE((((uintptr_t)ctx | (uintptr_t)cw) & 15) == 0, "ctx / cw must be 16-byte aligned");
  if (int rc = configure()) return rc;
  BwdArgs a{ctx, cw, lengths, attn, dwh, ld_dwh, dlogit_out, acts, c0, c1, d_h1_extra, d_c1, d_gates, d_c0, L, p, rng, call_off};
  const ChainLink link = vln_chain_link((cudaStream_t)stream, (unsigned int)B);
  VLN_CHECK_CUDA(vln_launch_linked(ctx_step_bwd_kernel, dim3(B), dim3(kThreads), sizeof(Smem), (cudaStream_t)stream, a, link));
  return 0;
}
