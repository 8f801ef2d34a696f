// Soft-dot attention over the instruction context [B, L, H] (fp32), forward and backward
// (SoftDotAttention.forward units.py:107-118: the bmm / masked softmax / bmm part).
//
// One CTA per episode: the episode's context tile (len x H fp32, <= 160 KB at L=80, H=512) is pulled
// into shared memory by ONE bulk-async copy (it is contiguous in [B,L,H]), so the dot products, the
// softmax and the weighted sum all run out of shared memory and the context is read from L2/HBM
// exactly once per call.  Everything stays inside the CTA — the first version split an episode over a
// 4-CTA cluster and paid two cluster barriers + DSMEM reads (~3 us) per call for it.
// Roofline class: L2/HBM — algorithmic bytes per episode-step: len*H*4 read, forward and backward.
// Backward can either accumulate d_context in place (read-modify-write of the whole tile: the
// autograd/module path) or just emit dlogit [B,L], leaving d_context to one batched GEMM over all
// steps at the end of the rollout (agent/fused.py).
#include "common.cuh"

namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxL = 96;

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// mode 0: forward. mode 1: backward.
__global__ void __launch_bounds__(kThreads, 1) ctx_attn_kernel(
    const float* __restrict__ context, const float* __restrict__ tgt, const int32_t* __restrict__ lengths,
    float* __restrict__ attn, float* __restrict__ weighted,            // fwd outputs (attn is an input in bwd)
    const float* __restrict__ d_weighted, const float* __restrict__ d_attn_ext, float* __restrict__ d_tgt,
    float* __restrict__ d_context, float* __restrict__ dlogit_out,       // bwd
    int mode, int L, int H, int ld_w, int staged, int ctx_ready) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* vec = reinterpret_cast<float*>(smem);           // [H]  tgt (fwd) / d_weighted (bwd)
  float* tg = vec + H;                                   // [H]  tgt (bwd only)
  float* sc = tg + H;                                    // [kMaxL] logits -> attn (fwd) / r_l -> dlogit (bwd)
  float* av = sc + kMaxL;                                // [kMaxL] saved attn (bwd)
  uint64_t* bar = reinterpret_cast<uint64_t*>(av + kMaxL);
  // [len][H] tile: in shared memory when it fits, else read in place (wide dense contexts, e.g. a
  // materialised [36, 2176] panorama)
  const float* tile = staged ? reinterpret_cast<const float*>(smem + (((size_t)(2 * H + 2 * kMaxL) * 4 + 16 + 127) & ~(size_t)127))
                             : context + (size_t)blockIdx.x * L * H;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  pdl_trigger();
  // ctx_ready: context / lengths (and, backward, the saved attention and tgt) were complete before the preceding kernel
  // started — true for every decoder step but the first — so the tile copy is requested BEFORE the programmatic
  // dependency wait and lands while the predecessor (the GEMM that produces tgt / d_weighted) is still finishing.
  if (!ctx_ready) pdl_wait();
  const int len = max(0, min(lengths[b], L));
  const uint32_t bytes = (uint32_t)len * (uint32_t)H * 4u;

  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
    if (staged && len > 0) {
      mbar_expect_tx(bar, bytes);
      bulk_g2s(const_cast<float*>(tile), context + (size_t)b * L * H, bytes, bar);
    }
  }
  if (mode == 1) {
    for (int i = tid; i < H; i += kThreads) tg[i] = tgt[(size_t)b * H + i];
    if (tid < L) av[tid] = tid < len ? attn[(size_t)b * L + tid] : 0.f;
  }
  if (ctx_ready) pdl_wait();
  const float* v_in = mode == 0 ? tgt + (size_t)b * H : d_weighted + (size_t)b * ld_w;
  for (int i = tid; i < H; i += kThreads) vec[i] = v_in[i];
  __syncthreads();                                       // barrier initialised; vec / tg / av in place
  if (staged && len > 0) mbar_wait(bar, 0);

  // ---- r_l = ctx_l . vec : one warp per row, 16-byte conflict-free shared loads ----
  const int nq = H / 128;                                // float4 chunks per lane
  for (int l = warp; l < len; l += kWarps) {
    const float4* row = reinterpret_cast<const float4*>(tile + (size_t)l * H);
    float a0 = 0.f, a1 = 0.f;
    for (int j = 0; j < nq; ++j) {
      const float4 x = row[j * 32 + lane], q = reinterpret_cast<const float4*>(vec)[j * 32 + lane];
      a0 = fmaf(x.x, q.x, a0); a1 = fmaf(x.y, q.y, a1);
      a0 = fmaf(x.z, q.z, a0); a1 = fmaf(x.w, q.w, a1);
    }
    const float s = warp_sum(a0 + a1);
    if (lane == 0) sc[l] = s;
  }
  __syncthreads();

  if (warp == 0) {
    float x[3], a[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int l = lane + 32 * k;
      x[k] = (l < len) ? sc[l] : -INFINITY;
    }
    if (mode == 0) {                                     // masked softmax (mask = position >= length)
      const float m = warp_max(fmaxf(x[0], fmaxf(x[1], x[2])));
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        a[k] = (lane + 32 * k < len) ? expf(x[k] - m) : 0.f;
        s += a[k];
      }
      const float inv = 1.0f / warp_sum(s);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int l = lane + 32 * k;
        if (l < L) {
          const float p = (l < len) ? a[k] * inv : 0.f;
          if (l < len) sc[l] = p;
          attn[(size_t)b * L + l] = p;
        }
      }
    } else {                                             // dlogit = attn * (r - attn . r)
      float dot = 0.f, p[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int l = lane + 32 * k;
        p[k] = (l < len) ? av[l] : 0.f;
        if (l < len && d_attn_ext) x[k] += d_attn_ext[(size_t)b * L + l];
        dot += (l < len) ? p[k] * x[k] : 0.f;
      }
      dot = warp_sum(dot);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int l = lane + 32 * k;
        const float d = (l < len) ? p[k] * (x[k] - dot) : 0.f;
        if (l < len) sc[l] = d;
        if (dlogit_out && l < L) dlogit_out[(size_t)b * L + l] = d;
      }
    }
  }
  __syncthreads();

  // ---- out[c] = sum_l sc[l] * ctx[l][c]  (weighted context / d_tgt) ----
  float* outp = mode == 0 ? weighted + (size_t)b * ld_w : d_tgt + (size_t)b * H;
  for (int c = tid; c < H; c += kThreads) {
    float a0 = 0.f, a1 = 0.f;
    int l = 0;
    for (; l + 1 < len; l += 2) {
      a0 = fmaf(sc[l], tile[(size_t)l * H + c], a0);
      a1 = fmaf(sc[l + 1], tile[(size_t)(l + 1) * H + c], a1);
    }
    if (l < len) a0 = fmaf(sc[l], tile[(size_t)l * H + c], a0);
    outp[c] = a0 + a1;
  }
  if (mode == 1 && d_context != nullptr) {               // d_ctx_l += attn_l * d_weighted + dlogit_l * tgt
    float* db = d_context + (size_t)b * L * H;
    const int hv = H / 4;
    for (int i = tid; i < len * hv; i += kThreads) {
      const int l = i / hv, c = i - l * hv;
      float4* p = reinterpret_cast<float4*>(db + (size_t)l * H) + c;
      float4 cur = *p;
      const float4 dw = reinterpret_cast<const float4*>(vec)[c], t4 = reinterpret_cast<const float4*>(tg)[c];
      const float a = av[l], dlg = sc[l];
      cur.x += a * dw.x + dlg * t4.x;
      cur.y += a * dw.y + dlg * t4.y;
      cur.z += a * dw.z + dlg * t4.z;
      cur.w += a * dw.w + dlg * t4.w;
      *p = cur;
    }
  }
}

int launch(const float* context, const float* tgt, const int32_t* lengths, float* attn, float* weighted,
           const float* d_weighted, const float* d_attn_ext, float* d_tgt, float* d_context, float* dlogit_out, int mode,
           int B, int L, int H, int ld_w, cudaStream_t stream, int ctx_ready = 0) {
  VLN_REQUIRE(L > 0 && L <= kMaxL, "L must be in 1..96");
  VLN_REQUIRE(H % 128 == 0 && H >= 128, "H must be a multiple of 128");
  VLN_REQUIRE(ld_w >= H, "row stride of weighted / d_weighted must be >= H");
  VLN_REQUIRE(((uintptr_t)context & 15) == 0, "context must be 16-byte aligned");
  const size_t head = (((size_t)(2 * H + 2 * kMaxL) * 4 + 16 + 127) & ~(size_t)127);
  const int staged = head + (size_t)L * H * 4 <= 220 * 1024;
  const size_t smem = head + (staged ? (size_t)L * H * 4 : 0);
  static size_t configured = 0;
  if (smem > configured) {
    VLN_CHECK_CUDA(cudaFuncSetAttribute(ctx_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  VLN_CHECK_CUDA(vln_launch_chain(ctx_attn_kernel, dim3(B), dim3(kThreads), smem, stream, context, tgt, lengths, attn,
                                  weighted, d_weighted, d_attn_ext, d_tgt, d_context, dlogit_out, mode, L, H, ld_w, staged,
                                  ctx_ready && staged ? 1 : 0));
  return 0;
}

}  // namespace

extern "C" int vln_ctx_attn_fwd_ld(const float* context, const float* tgt, const int32_t* lengths, float* attn,
                                   float* weighted, int ld_weighted, int B, int L, int H, int ctx_ready, void* stream) {
  VLN_REQUIRE(context && tgt && lengths && attn && weighted && B > 0, "bad arguments");
  return launch(context, tgt, lengths, attn, weighted, nullptr, nullptr, nullptr, nullptr, nullptr, 0, B, L, H,
                ld_weighted, (cudaStream_t)stream, ctx_ready);
}

extern "C" int vln_ctx_attn_fwd(const float* context, const float* tgt, const int32_t* lengths, float* attn,
                                float* weighted, int B, int L, int H, void* stream) {
  return vln_ctx_attn_fwd_ld(context, tgt, lengths, attn, weighted, H, B, L, H, 0, stream);
}

extern "C" int vln_ctx_attn_bwd_ld(const float* context, const float* tgt, const int32_t* lengths, const float* attn,
                                   const float* d_weighted, int ld_d_weighted, const float* d_attn_ext, float* d_tgt,
                                   float* d_context, float* dlogit_out, int B, int L, int H, int ctx_ready, void* stream) {
  VLN_REQUIRE(context && tgt && lengths && attn && d_weighted && d_tgt && B > 0, "bad arguments");
  return launch(context, tgt, lengths, const_cast<float*>(attn), nullptr, d_weighted, d_attn_ext, d_tgt, d_context,
                dlogit_out, 1, B, L, H, ld_d_weighted, (cudaStream_t)stream, ctx_ready);
}

extern "C" int vln_ctx_attn_bwd(const float* context, const float* tgt, const int32_t* lengths, const float* attn,
                                const float* d_weighted, const float* d_attn_ext, float* d_tgt, float* d_context,
                                int B, int L, int H, void* stream) {
  return vln_ctx_attn_bwd_ld(context, tgt, lengths, attn, d_weighted, H, d_attn_ext, d_tgt, d_context, nullptr, B, L, H,
                             0, stream);
}
