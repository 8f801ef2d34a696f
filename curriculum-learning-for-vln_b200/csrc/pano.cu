// Feature-table kernels: bit-exact gathers, the fused TMA gather + 36-view soft-dot attention
// (forward and backward share one kernel), candidate logits forward/backward.
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
// K3: fused gather + soft-dot attention over the panorama, forward (mode 0) / backward (mode 1).
//
// grid = (S, B), cluster = (S,1,1).  CTA `rank` of episode b owns image features
// [rank*FS, (rank+1)*FS), FS = 2048/S: one TMA box per 256 columns lands the [36 x FS] bf16 slice
// in shared memory (the only HBM read of the tile).  Per-view partial dot products are summed
// across the cluster through distributed shared memory (36 floats per CTA), every CTA then holds
// the full softmax (or its Jacobian-vector product) and produces its own FS output columns.
// rank 0 also carries the 128 angle dimensions, which are only 4 distinct values per view.
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) pano_attn_kernel(const __grid_constant__ CUtensorMap tmap,
                                                             const int32_t* __restrict__ vp,
                                                             const int32_t* __restrict__ view,
                                                             const float* __restrict__ loc4,
                                                             const float* __restrict__ vec,
                                                             float* __restrict__ attn_io, float* __restrict__ out,
                                                             int mode, float drop_p, const uint64_t* __restrict__ rng, uint64_t call_off,
                                                             int FS) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rank = (int)cluster_ctarank();
  const int S = (int)cluster_nctarank();
  const int b = blockIdx.y;
  const int nbox = FS / kBoxCols;

  __nv_bfloat16* tile = reinterpret_cast<__nv_bfloat16*>(smem);
  float* vsl = reinterpret_cast<float*>(smem + (size_t)nbox * kBoxBytes);   // [FS] slice of q / d(out)
  float* part = vsl + FS;                                                   // [36] (+4 pad)
  float* sm = part + 40;                                                    // [36] (+4 pad)
  float* loc = sm + 40;                                                     // [36*4]
  float* qa = loc + 144;                                                    // [4]
  uint64_t* bar = reinterpret_cast<uint64_t*>(qa + 4);

  const int g = vp[b];
  const int cur_view = view[b];
  if (tid == 0) {
    tma_prefetch_desc(&tmap);
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(bar, (uint32_t)(nbox * kBoxBytes));
    for (int i = 0; i < nbox; ++i)
      tma_load_2d(tile + (size_t)i * VLN_V * kBoxCols, &tmap, bar, rank * FS + i * kBoxCols, g * VLN_V);
  }
  // overlap with the TMA: stage the vector slice, the angle table row and the angle-group sums
  const float* vrow = vec + (size_t)b * VLN_F;
  for (int i = tid; i < FS; i += kThreads) vsl[i] = vrow[rank * FS + i];
  for (int i = tid; i < 144; i += kThreads) loc[i] = loc4[(size_t)cur_view * 144 + i];
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float s = warp_sum(vrow[VLN_IMG + 32 * k + lane]);
      if (lane == 0) qa[k] = s;
    }
  }
  mbar_wait(bar, 0);

  float scale = 1.0f;
  if (drop_p > 0.f) {   // policy.py:226-231 — feature dropout on the 2048 image dims only
    scale = 1.0f / (1.0f - drop_p);
    const uint64_t seed = rng[0], offset = rng[1] + call_off;
    const uint32_t thr = drop_threshold(drop_p);
    const int nvec = nbox * VLN_V * 32;
    uint4* t4 = reinterpret_cast<uint4*>(tile);
    for (int i = tid; i < nvec; i += kThreads) {
      const int box = i / (VLN_V * 32), r = i - box * (VLN_V * 32), v = r >> 5, j = r & 31;
      const uint64_t e = (((uint64_t)b * VLN_V + v) * VLN_IMG + (uint64_t)(rank * FS + box * kBoxCols + j * 8)) >> 3;
      t4[i] = apply_keep(t4[i], philox8(seed, offset, e), thr);
    }
  }
  __syncthreads();

  // phase 1: partial dot products, one warp per view
  for (int v = warp; v < VLN_V; v += kThreads / 32) {
    float acc = 0.f;
    for (int box = 0; box < nbox; ++box) {
      const uint4 x = reinterpret_cast<const uint4*>(tile)[(box * VLN_V + v) * 32 + lane];
      const float4 q0 = reinterpret_cast<const float4*>(vsl)[(box * kBoxCols + lane * 8) / 4];
      const float4 q1 = reinterpret_cast<const float4*>(vsl)[(box * kBoxCols + lane * 8) / 4 + 1];
      acc += bf16lo(x.x) * q0.x + bf16hi(x.x) * q0.y + bf16lo(x.y) * q0.z + bf16hi(x.y) * q0.w +
             bf16lo(x.z) * q1.x + bf16hi(x.z) * q1.y + bf16lo(x.w) * q1.z + bf16hi(x.w) * q1.w;
    }
    acc = warp_sum(acc) * scale;
    if (lane == 0) {
      if (rank == 0)
        acc += loc[v * 4] * qa[0] + loc[v * 4 + 1] * qa[1] + loc[v * 4 + 2] * qa[2] + loc[v * 4 + 3] * qa[3];
      part[v] = acc;
    }
  }
  cluster_arrive();
  cluster_wait();
  if (tid < VLN_V) {
    float tot = 0.f;
    for (int r = 0; r < S; ++r) tot += dsmem_ld_f32(part + tid, (uint32_t)r);
    sm[tid] = tot;
  }
  cluster_arrive();          // peers may retire their `part` once everyone has read it (wait at exit)
  __syncthreads();

  if (warp == 0) {
    const bool has1 = lane + 32 < VLN_V;
    const float x0 = sm[lane], x1 = has1 ? sm[lane + 32] : -INFINITY;
    float a0, a1;
    if (mode == 0) {
      const float m = warp_max(fmaxf(x0, x1));
      const float e0 = expf(x0 - m), e1 = has1 ? expf(x1 - m) : 0.f;
      const float inv = 1.0f / warp_sum(e0 + e1);
      a0 = e0 * inv;
      a1 = e1 * inv;
      if (rank == 0) {
        attn_io[(size_t)b * VLN_V + lane] = a0;
        if (has1) attn_io[(size_t)b * VLN_V + lane + 32] = a1;
      }
    } else {
      const float p0 = attn_io[(size_t)b * VLN_V + lane], p1 = has1 ? attn_io[(size_t)b * VLN_V + lane + 32] : 0.f;
      const float dot = warp_sum(p0 * x0 + (has1 ? p1 * x1 : 0.f));
      a0 = p0 * (x0 - dot);
      a1 = has1 ? p1 * (x1 - dot) : 0.f;
    }
    sm[lane] = a0;
    if (has1) sm[lane + 32] = a1;
  }
  __syncthreads();

  // phase 2: this CTA's output columns, two features per thread
  float* orow = out + (size_t)b * VLN_F;
  for (int p = tid; p < FS / 2; p += kThreads) {
    const int f = 2 * p, box = f / kBoxCols, col = f - box * kBoxCols;
    const uint32_t* colp = reinterpret_cast<const uint32_t*>(tile + (size_t)box * VLN_V * kBoxCols + col);
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 6
    for (int v = 0; v < VLN_V; ++v) {
      const uint32_t w = colp[v * (kBoxCols / 2)];
      acc0 += sm[v] * bf16lo(w);
      acc1 += sm[v] * bf16hi(w);
    }
    reinterpret_cast<float2*>(orow + rank * FS)[p] = make_float2(acc0 * scale, acc1 * scale);
  }
  if (rank == 0 && tid < VLN_ANG) {
    const int k = tid >> 5;
    float acc = 0.f;
    for (int v = 0; v < VLN_V; ++v) acc += sm[v] * loc[v * 4 + k];
    orow[VLN_IMG + tid] = acc;
  }
  cluster_wait();
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

// d_tgt[b, f] = sum_j dlogits[b, j] * x~c[b, j, f];  one 16-byte vector (8 features) per thread.
__global__ void __launch_bounds__(kThreads) cand_logits_bwd_kernel(
    const __nv_bfloat16* __restrict__ table, const int32_t* __restrict__ vp, const int32_t* __restrict__ view,
    const int32_t* __restrict__ cand_view, const float* __restrict__ cand_ang4, const int32_t* __restrict__ n_cand,
    const float* __restrict__ dlogits, float* __restrict__ d_tgt, float* __restrict__ d_bias, float drop_p,
    const uint64_t* __restrict__ rng, uint64_t call_off) {
  __shared__ float dl[VLN_NSLOT];
  __shared__ int cvs[VLN_NSLOT];
  const int b = blockIdx.x, t = threadIdx.x;
  const int g = vp[b];
  const int n = n_cand[g];
  if (t < VLN_NSLOT) {
    dl[t] = (t <= n) ? dlogits[(size_t)b * VLN_NSLOT + t] : 0.f;
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
      const uint64_t e = (((uint64_t)b * VLN_NSLOT + j) * VLN_IMG + (uint64_t)t * 8) >> 3;
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

size_t pano_smem_bytes(int FS) { return (size_t)(FS / kBoxCols) * kBoxBytes + (size_t)FS * 4 + (40 + 40 + 144 + 4) * 4 + 16; }

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

extern "C" int vln_pano_attn(const vln_ctx* ctx, const int32_t* vp, const int32_t* view, const float* loc4,
                             const float* vec, float* attn_io, float* out, int B, int mode, float drop_p,
                             const uint64_t* rng, uint64_t call_off, int split, void* stream) {
  VLN_REQUIRE(ctx && vp && view && loc4 && vec && attn_io && out && B > 0, "bad arguments");
  VLN_REQUIRE(split == 1 || split == 2 || split == 4 || split == 8, "split must be 1, 2, 4 or 8");
  VLN_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (forward) or 1 (backward)");
  VLN_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "drop_p out of range");
  VLN_REQUIRE(drop_p == 0.f || rng, "dropout needs an rng state");
  const int FS = VLN_IMG / split;
  const size_t smem = pano_smem_bytes(FS);
  static size_t configured = 0;
  if (smem > configured) {
    VLN_CHECK_CUDA(cudaFuncSetAttribute(pano_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(split, B);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = split;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  VLN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, pano_attn_kernel, ctx->tmap_tile, vp, view, loc4, vec, attn_io, out, mode,
                                    drop_p, rng, call_off, FS));
  return 0;
}

extern "C" int vln_cand_logits_fwd(const vln_ctx* ctx, const int32_t* vp, const int32_t* view,
                                   const int32_t* cand_view, const float* cand_ang4, const int32_t* n_cand,
                                   const float* tgt, const float* bias, float* logits, int B, float drop_p,
                                   const uint64_t* rng, uint64_t call_off, void* stream) {
  VLN_REQUIRE(ctx && vp && view && cand_view && cand_ang4 && n_cand && tgt && logits && B > 0, "bad arguments");
  cand_logits_fwd_kernel<<<B, 512, 0, (cudaStream_t)stream>>>(ctx->table, vp, view, cand_view, cand_ang4, n_cand, tgt,
                                                             bias, logits, drop_p, rng, call_off);
  VLN_LAUNCH_OK();
  return 0;
}

extern "C" int vln_cand_logits_bwd(const vln_ctx* ctx, const int32_t* vp, const int32_t* view,
                                   const int32_t* cand_view, const float* cand_ang4, const int32_t* n_cand,
                                   const float* dlogits, float* d_tgt, float* d_bias, int B, float drop_p,
                                   const uint64_t* rng, uint64_t call_off, void* stream) {
  VLN_REQUIRE(ctx && vp && view && cand_view && cand_ang4 && n_cand && dlogits && d_tgt && B > 0, "bad arguments");
  cand_logits_bwd_kernel<<<B, kThreads, 0, (cudaStream_t)stream>>>(ctx->table, vp, view, cand_view, cand_ang4, n_cand,
                                                                  dlogits, d_tgt, d_bias, drop_p, rng, call_off);
  VLN_LAUNCH_OK();
  return 0;
}
