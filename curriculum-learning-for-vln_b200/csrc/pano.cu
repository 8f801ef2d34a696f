// Feature-table kernels: bit-exact gathers and candidate logits forward/backward (the fused
// gather + 36-view attention lives in pano_attn.cu).
//
// Roofline class: HBM.  Algorithmic bytes per episode-step: 36*2048*2 = 147 456 B (pano tile),
// n_cand*4096 B (candidate rows).  See DESIGN.md §Kernels.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kBoxCols = 256;                          // TMA box: 256 bf16 columns x 36 rows = 18 432 B
constexpr int kBoxBytes = VLN_V * kBoxCols * 2;

__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// zero the dropped bf16 lanes of one 16-byte vector (8 features) in place
__device__ __forceinline__ uint4 apply_keep(uint4 x, const Philox8& r, uint32_t thr) {
  uint32_t* w = reinterpret_cast<uint32_t*>(&x);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t m = (philox_keep(r, 2 * i, thr) ? 0x0000FFFFu : 0u) | (philox_keep(r, 2 * i + 1, thr) ? 0xFFFF0000u : 0u);
    w[i] &= m;
  }
  return x;
}

// -------------------------------------------------------------------------------------------
// K1: out[b, v, :] = concat(table[vp[b], v, :] as fp32, loc4[view[b], v, k] repeated 32x)
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) gather_pano_kernel(const __nv_bfloat16* __restrict__ table,
                                                               const int32_t* __restrict__ vp,
                                                               const int32_t* __restrict__ view,
                                                               const float* __restrict__ loc4,
                                                               float* __restrict__ out) {
  const int v = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
  const uint4* src = reinterpret_cast<const uint4*>(table + ((size_t)vp[b] * VLN_V + v) * VLN_IMG);
  float* dst = out + ((size_t)b * VLN_V + v) * VLN_F;
  uint4 x = __ldg(src + t);
  float4 lo = make_float4(bf16lo(x.x), bf16hi(x.x), bf16lo(x.y), bf16hi(x.y));
  float4 hi = make_float4(bf16lo(x.z), bf16hi(x.z), bf16lo(x.w), bf16hi(x.w));
  reinterpret_cast<float4*>(dst)[2 * t] = lo;
  reinterpret_cast<float4*>(dst)[2 * t + 1] = hi;
  if (t < VLN_ANG) dst[VLN_IMG + t] = loc4[((size_t)view[b] * VLN_V + v) * 4 + (t >> 5)];
}

// -------------------------------------------------------------------------------------------
// K2: candidate rows, zero padded; row n_cand (END) and beyond are zero.
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) gather_cand_kernel(
    const __nv_bfloat16* __restrict__ table, const int32_t* __restrict__ vp, const int32_t* __restrict__ view,
    const int32_t* __restrict__ cand_view, const float* __restrict__ cand_ang4, const int32_t* __restrict__ n_cand,
    float* __restrict__ out, int32_t* __restrict__ out_len, int C) {
  const int j = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
  const int g = vp[b];
  const int n = n_cand[g];
  float* dst = out + ((size_t)b * C + j) * VLN_F;
  if (j == 0 && t == 0 && out_len) out_len[b] = n + 1;
  if (j < n) {
    const int cv = cand_view[(size_t)g * VLN_CMAX + j];
    const uint4* src = reinterpret_cast<const uint4*>(table + ((size_t)g * VLN_V + cv) * VLN_IMG);
    uint4 x = __ldg(src + t);
    reinterpret_cast<float4*>(dst)[2 * t] = make_float4(bf16lo(x.x), bf16hi(x.x), bf16lo(x.y), bf16hi(x.y));
    reinterpret_cast<float4*>(dst)[2 * t + 1] = make_float4(bf16lo(x.z), bf16hi(x.z), bf16lo(x.w), bf16hi(x.w));
    if (t < VLN_ANG)
      dst[VLN_IMG + t] = cand_ang4[(((size_t)g * VLN_CMAX + j) * 12 + (view[b] % 12)) * 4 + (t >> 5)];
  } else {
    reinterpret_cast<float4*>(dst)[2 * t] = make_float4(0.f, 0.f, 0.f, 0.f);
    reinterpret_cast<float4*>(dst)[2 * t + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < VLN_ANG) dst[VLN_IMG + t] = 0.f;
  }
}

// a_t_prev = cands[b, max(action,0)] (follower.py:164): one candidate row per episode.
__global__ void __launch_bounds__(kThreads) gather_action_kernel(
    const __nv_bfloat16* __restrict__ table, const int32_t* __restrict__ vp, const int32_t* __restrict__ view,
    const int32_t* __restrict__ action, const uint8_t* __restrict__ ended, const int32_t* __restrict__ cand_view,
    const float* __restrict__ cand_ang4, const int32_t* __restrict__ n_cand, float* __restrict__ out) {
  const int b = blockIdx.x, t = threadIdx.x;
  const int g = vp[b];
  const int a = action[b];
  const int j = ((ended && ended[b]) || a < 0 || a >= n_cand[g]) ? 0 : a;
  float* dst = out + (size_t)b * VLN_F;
  if (j < n_cand[g]) {
    const int cv = cand_view[(size_t)g * VLN_CMAX + j];
    const uint4 x = __ldg(reinterpret_cast<const uint4*>(table + ((size_t)g * VLN_V + cv) * VLN_IMG) + t);
    reinterpret_cast<float4*>(dst)[2 * t] = make_float4(bf16lo(x.x), bf16hi(x.x), bf16lo(x.y), bf16hi(x.y));
    reinterpret_cast<float4*>(dst)[2 * t + 1] = make_float4(bf16lo(x.z), bf16hi(x.z), bf16lo(x.w), bf16hi(x.w));
    if (t < VLN_ANG)
      dst[VLN_IMG + t] = cand_ang4[(((size_t)g * VLN_CMAX + j) * 12 + (view[b] % 12)) * 4 + (t >> 5)];
  } else {
    reinterpret_cast<float4*>(dst)[2 * t] = make_float4(0.f, 0.f, 0.f, 0.f);
    reinterpret_cast<float4*>(dst)[2 * t + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < VLN_ANG) dst[VLN_IMG + t] = 0.f;
  }
}

// -------------------------------------------------------------------------------------------
// K6: candidate logits.  grid = B, 16 warps: warp j scores candidate slot j.
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) cand_logits_fwd_kernel(
    const __nv_bfloat16* __restrict__ table, const int32_t* __restrict__ vp, const int32_t* __restrict__ view,
    const int32_t* __restrict__ cand_view, const float* __restrict__ cand_ang4, const int32_t* __restrict__ n_cand,
    const float* __restrict__ tgt, const float* __restrict__ bias, float* __restrict__ logits, float drop_p,
    const uint64_t* __restrict__ rng, uint64_t call_off) {
  __shared__ __align__(16) float ts[VLN_F];
  __shared__ float ta[4];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, j = tid >> 5;
  pdl_trigger();
  pdl_wait();
  const int g = vp[b];
  const int n = n_cand[g];
  for (int i = tid; i < VLN_F; i += 512) ts[i] = tgt[(size_t)b * VLN_F + i];
  __syncthreads();
  if (j < 4) {
    float s = warp_sum(ts[VLN_IMG + 32 * j + lane]);
    if (lane == 0) ta[j] = s;
  }
  __syncthreads();
  const float bb = bias ? bias[b] : 0.f;
  float res;
  if (j < n) {
    const int cv = cand_view[(size_t)g * VLN_CMAX + j];
    const uint4* src = reinterpret_cast<const uint4*>(table + ((size_t)g * VLN_V + cv) * VLN_IMG);
    const uint32_t thr = drop_threshold(drop_p);
    uint64_t seed = 0, offset = 0;
    if (drop_p > 0.f) { seed = rng[0]; offset = rng[1] + call_off; }
    float acc = 0.f;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int vi = it * 32 + lane;
      uint4 x = __ldg(src + vi);
      if (drop_p > 0.f) {
        const uint64_t e = (((uint64_t)b * VLN_NSLOT + j) * VLN_IMG + (uint64_t)vi * 8) >> 3;
        x = apply_keep(x, philox8(seed, offset, e), thr);
      }
      const float4 q0 = reinterpret_cast<const float4*>(ts)[vi * 2], q1 = reinterpret_cast<const float4*>(ts)[vi * 2 + 1];
      acc += bf16lo(x.x) * q0.x + bf16hi(x.x) * q0.y + bf16lo(x.y) * q0.z + bf16hi(x.y) * q0.w +
             bf16lo(x.z) * q1.x + bf16hi(x.z) * q1.y + bf16lo(x.w) * q1.z + bf16hi(x.w) * q1.w;
    }
    acc = warp_sum(acc);
    if (drop_p > 0.f) acc *= 1.0f / (1.0f - drop_p);
    const float* ang = cand_ang4 + (((size_t)g * VLN_CMAX + j) * 12 + (view[b] % 12)) * 4;
    res = acc + ang[0] * ta[0] + ang[1] * ta[1] + ang[2] * ta[2] + ang[3] * ta[3] + bb;
  } else if (j == n) {
    res = bb;                                    // END slot: all-zero feature row (base.py:152-153)
  } else {
    res = -INFINITY;                             // length2mask + masked_fill_(-inf)
  }
  if (lane == 0) logits[(size_t)b * VLN_NSLOT + j] = res;
}

struct PolicyGrad {                                  // optional: compute dlogits in place of reading them
  const float* probs; const int32_t* target; const int32_t* action; const float* entropy;
  const float* g_ce; const float* g_logp; const float* g_ent;
};

// d_tgt[b, f] = sum_j dlogits[b, j] * x~c[b, j, f];  one 16-byte vector (8 features) per thread.
__global__ void __launch_bounds__(kThreads) cand_logits_bwd_kernel(
    const __nv_bfloat16* __restrict__ table, const int32_t* __restrict__ vp, const int32_t* __restrict__ view,
    const int32_t* __restrict__ cand_view, const float* __restrict__ cand_ang4, const int32_t* __restrict__ n_cand,
    const float* __restrict__ dlogits, float* __restrict__ d_tgt, float* __restrict__ d_bias, float drop_p,
    const uint64_t* __restrict__ rng, uint64_t call_off, PolicyGrad pg, int B_step, uint64_t off_stride) {
  __shared__ float dl[VLN_NSLOT];
  __shared__ int cvs[VLN_NSLOT];
  // rows may stack several decoder steps ([n_steps, B_step]): row b of step s draws its mask from stream
  // call_off + s*off_stride with the element indices of a single [B_step,16,2048] tensor
  const int b = blockIdx.x, t = threadIdx.x;
  const int bl = b % B_step;
  call_off += (uint64_t)(b / B_step) * off_stride;
  const int g = vp[b];
  const int n = n_cand[g];
  if (t < VLN_NSLOT) {
    float d = 0.f;
    if (pg.probs) {                                  // dlogits from the action head's saved state (vln_policy_bwd)
      const float p = pg.probs[(size_t)b * VLN_NSLOT + t];
      if (pg.g_ce && pg.target[b] >= 0) d += pg.g_ce[b] * (p - (t == pg.target[b] ? 1.f : 0.f));
      if (pg.g_logp && pg.action[b] >= 0) d += pg.g_logp[b] * ((t == pg.action[b] ? 1.f : 0.f) - p);
      if (pg.g_ent && p > 0.f) d -= pg.g_ent[b] * p * (logf(p) + pg.entropy[b]);
    } else {
      d = dlogits[(size_t)b * VLN_NSLOT + t];
    }
    dl[t] = (t <= n) ? d : 0.f;
    cvs[t] = (t < n) ? cand_view[(size_t)g * VLN_CMAX + t] : 0;
  }
  __syncthreads();
  const uint32_t thr = drop_threshold(drop_p);
  uint64_t seed = 0, offset = 0;
  if (drop_p > 0.f) { seed = rng[0]; offset = rng[1] + call_off; }
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int j = 0; j < n; ++j) {
    uint4 x = __ldg(reinterpret_cast<const uint4*>(table + ((size_t)g * VLN_V + cvs[j]) * VLN_IMG) + t);
    if (drop_p > 0.f) {
      const uint64_t e = (((uint64_t)bl * VLN_NSLOT + j) * VLN_IMG + (uint64_t)t * 8) >> 3;
      x = apply_keep(x, philox8(seed, offset, e), thr);
    }
    const float d = dl[j];
    acc[0] += d * bf16lo(x.x); acc[1] += d * bf16hi(x.x); acc[2] += d * bf16lo(x.y); acc[3] += d * bf16hi(x.y);
    acc[4] += d * bf16lo(x.z); acc[5] += d * bf16hi(x.z); acc[6] += d * bf16lo(x.w); acc[7] += d * bf16hi(x.w);
  }
  const float sc = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
  float* dst = d_tgt + (size_t)b * VLN_F;
  reinterpret_cast<float4*>(dst)[2 * t] = make_float4(acc[0] * sc, acc[1] * sc, acc[2] * sc, acc[3] * sc);
  reinterpret_cast<float4*>(dst)[2 * t + 1] = make_float4(acc[4] * sc, acc[5] * sc, acc[6] * sc, acc[7] * sc);
  if (t < VLN_ANG) {
    const int k = t >> 5, vm = view[b] % 12;
    float a = 0.f;
    for (int j = 0; j < n; ++j) a += dl[j] * cand_ang4[(((size_t)g * VLN_CMAX + j) * 12 + vm) * 4 + k];
    dst[VLN_IMG + t] = a;
  }
  if (d_bias && t == 0) {
    float s = 0.f;
    for (int j = 0; j <= n; ++j) s += dl[j];
    d_bias[b] = s;
  }
}

}  // namespace

extern "C" int vln_gather_pano(const vln_ctx* ctx, const int32_t* vp, const int32_t* view, const float* loc4,
                               float* out, int B, void* stream) {
  VLN_REQUIRE(ctx && vp && view && loc4 && out && B > 0, "bad arguments");
  gather_pano_kernel<<<dim3(VLN_V, B), kThreads, 0, (cudaStream_t)stream>>>(ctx->table, vp, view, loc4, out);
  VLN_LAUNCH_OK();
  return 0;
}

extern "C" int vln_gather_cand(const vln_ctx* ctx, const int32_t* vp, const int32_t* view, const int32_t* cand_view,
                               const float* cand_ang4, const int32_t* n_cand, float* out, int32_t* out_len, int B,
                               int C, void* stream) {
  VLN_REQUIRE(ctx && vp && view && cand_view && cand_ang4 && n_cand && out && B > 0 && C > 0, "bad arguments");
  gather_cand_kernel<<<dim3(C, B), kThreads, 0, (cudaStream_t)stream>>>(ctx->table, vp, view, cand_view, cand_ang4,
                                                                        n_cand, out, out_len, C);
  VLN_LAUNCH_OK();
  return 0;
}

extern "C" int vln_gather_action_feat(const vln_ctx* ctx, const int32_t* vp, const int32_t* view,
                                      const int32_t* action, const uint8_t* ended, const int32_t* cand_view,
                                      const float* cand_ang4, const int32_t* n_cand, float* out, int B,
                                      void* stream) {
  VLN_REQUIRE(ctx && vp && view && action && cand_view && cand_ang4 && n_cand && out && B > 0, "bad arguments");
  gather_action_kernel<<<B, kThreads, 0, (cudaStream_t)stream>>>(ctx->table, vp, view, action, ended, cand_view,
                                                                cand_ang4, n_cand, out);
  VLN_LAUNCH_OK();
  return 0;
}

extern "C" int vln_cand_logits_fwd(const vln_ctx* ctx, const int32_t* vp, const int32_t* view,
                                   const int32_t* cand_view, const float* cand_ang4, const int32_t* n_cand,
                                   const float* tgt, const float* bias, float* logits, int B, float drop_p,
                                   const uint64_t* rng, uint64_t call_off, void* stream) {
  VLN_REQUIRE(ctx && vp && view && cand_view && cand_ang4 && n_cand && tgt && logits && B > 0, "bad arguments");
  VLN_CHECK_CUDA(vln_launch_chain(cand_logits_fwd_kernel, dim3(B), dim3(512), 0, (cudaStream_t)stream, ctx->table, vp, view,
                                  cand_view, cand_ang4, n_cand, tgt, bias, logits, drop_p, rng, call_off));
  return 0;
}

extern "C" int vln_cand_logits_bwd(const vln_ctx* ctx, const int32_t* vp, const int32_t* view,
                                   const int32_t* cand_view, const float* cand_ang4, const int32_t* n_cand,
                                   const float* dlogits, float* d_tgt, float* d_bias, int B, float drop_p,
                                   const uint64_t* rng, uint64_t call_off, void* stream) {
  VLN_REQUIRE(ctx && vp && view && cand_view && cand_ang4 && n_cand && dlogits && d_tgt && B > 0, "bad arguments");
  cand_logits_bwd_kernel<<<B, kThreads, 0, (cudaStream_t)stream>>>(ctx->table, vp, view, cand_view, cand_ang4, n_cand,
                                                                  dlogits, d_tgt, d_bias, drop_p, rng, call_off,
                                                                  PolicyGrad{}, B, 0);
  VLN_LAUNCH_OK();
  return 0;
}

extern "C" int vln_cand_logits_bwd_policy(const vln_ctx* ctx, const int32_t* vp, const int32_t* view,
                                          const int32_t* cand_view, const float* cand_ang4, const int32_t* n_cand,
                                          const float* probs, const int32_t* target, const int32_t* action,
                                          const float* entropy, const float* g_ce, const float* g_logp,
                                          const float* g_ent, float* d_tgt, int B, int n_steps, float drop_p,
                                          const uint64_t* rng, uint64_t call_off, uint64_t off_stride, void* stream) {
  VLN_REQUIRE(ctx && vp && view && cand_view && cand_ang4 && n_cand && probs && d_tgt && B > 0 && n_steps > 0,
              "bad arguments");
  VLN_REQUIRE((!g_ce || target) && (!g_logp || action) && (!g_ent || entropy), "missing saved action-head state");
  cand_logits_bwd_kernel<<<B * n_steps, kThreads, 0, (cudaStream_t)stream>>>(
      ctx->table, vp, view, cand_view, cand_ang4, n_cand, nullptr, d_tgt, nullptr, drop_p, rng, call_off,
      PolicyGrad{probs, target, action, entropy, g_ce, g_logp, g_ent}, B, off_stride);
  VLN_LAUNCH_OK();
  return 0;
}
