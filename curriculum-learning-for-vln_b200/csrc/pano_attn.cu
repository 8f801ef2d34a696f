// Fused gather + feature dropout + soft-dot attention over the 36-view panorama
// (SoftDotAttention.forward units.py:107-118 on `_feature_variable` base.py:141-147 with
// EnvDropDecoder's env_drop policy.py:226-231), forward and backward in one kernel.
//
// Roofline class: HBM.  Algorithmic bytes: 36 x 2048 x 2 = 147 456 B per episode-step, read ONCE.
//
// v3: the whole panorama of one episode (36 rows x 4 096 B, + 36 x 256 B of packed keep-bits) sits in
// shared memory, and the unit is processed in two phases by all 12 warps:
//   phase 1  warp w owns rows w, w+12, w+24: waits for the row's bulk-async copy (one mbarrier per row),
//            applies the keep-mask (and writes the masked row back, so phase 2 never touches the
//            mask), dots it with the register-resident query -> logit[v];
//   softmax  one warp: a = softmax(logit) (forward) or c_v = a_v (r_v - sum_u a_u r_u) (backward);
//   phase 2  column-parallel: thread group g (128 threads) owns rows 12g..12g+11, every thread two
//            16-byte column chunks: acc += w_v * row_v[chunk] — no cross-warp merge of 2048-wide
//            accumulators (v2 spent a third of its time merging six of them through shared memory).
// The kernel is persistent (CTA c takes episodes c, c+grid, ...).  Rows are released in four groups
// of nine as phase 2 finishes with them, and the NEXT episode's rows are requested into the freed
// slots at once, so the copy engine keeps streaming while phase 2 and the next phase 1 run.
// The 128 angle dimensions are 4 distinct values per view (misc.py:286-293): they ride along as 4
// scalars per row (loc4) against 4 group sums of the query.
//
//   forward :  logit_v = x~_v . q          a = softmax(logit)      out = sum_v a_v x~_v
//   backward:  r_v     = x~_v . d_out      c_v = a_v (r_v - a.r)   dq  = sum_v c_v x~_v
// with x~ = [ dropout(table row) | angle embedding ].
#include <cstdlib>

#include "common.cuh"

// optional phase stamps (VLN_PANO_STAMPS=1): CTA 0 / thread 0 records clock64 at phase boundaries of its first unit
__device__ unsigned long long g_pano_stamps[16];
#define PSTAMP(i)                                                                          \
  do {                                                                                     \
    if (dbg && blockIdx.x == 0 && tid == 0 && it == 0) g_pano_stamps[i] = (unsigned long long)clock64(); \
  } while (0)

namespace {

constexpr int kWarps = 12;
constexpr int kThreads = kWarps * 32;                    // 384
constexpr int kRowBytes = VLN_IMG * 2;                   // 4 096
constexpr int kMaskBytes = VLN_IMG / 8;                  // 256: packed keep-bits of one row (vln_feature_mask_bits)
constexpr int kGroups = 3;                               // phase-2 thread groups (128 threads, 12 rows each)
constexpr int kRowsPerGroup = VLN_V / kGroups;           // 12
constexpr int kRel = 4;                                  // release points per unit (3 rows of every group each)

__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct Smem {
  uint8_t rows[VLN_V * kRowBytes];          // the episode's panorama, bf16
  uint8_t mask[VLN_V * kMaskBytes];         // packed keep-bits (only with pre-generated masks)
  float qbuf[2][VLN_F];                     // query / d_out row of this and the next unit (cp.async double buffer)
  float locbuf[2][VLN_V * 4];               // loc4[cur_view] rows
  float attbuf[2][40];                      // saved attention (backward)
  float logit[40];
  float wv[40];                             // softmax weights (forward) / c_v (backward)
  float part[(kGroups - 1) * 128 * 16];     // phase-2 partial sums of groups 1, 2
  uint64_t full[VLN_V];
};

__global__ void __launch_bounds__(kThreads, 1)
pano_attn_kernel(const __nv_bfloat16* __restrict__ table, const int32_t* __restrict__ vp,
                 const int32_t* __restrict__ view, const float* __restrict__ loc4, const float* __restrict__ vec,
                 float* __restrict__ attn_io, float* __restrict__ out, int B, int mode, float drop_p,
                 const uint64_t* __restrict__ rng, uint64_t call_off, int ld_vec, int ld_out,
                 const uint8_t* __restrict__ mask_bits, int dbg) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  int it = 0;
  PSTAMP(0);
  pdl_trigger();
  if (tid == 0) {
    for (int r = 0; r < VLN_V; ++r) mbar_init(&sm.full[r], 1);
    fence_mbar_init();
  }
  __syncthreads();
  pdl_wait();                                            // viewpoints, query and mask bits come from predecessors
  PSTAMP(1);

  // request row r of episode `ep` (and its keep-bits) into its slot
  auto request_row = [&](int ep, int r) {
    const uint32_t bytes = mask_bits ? kRowBytes + kMaskBytes : kRowBytes;
    mbar_expect_tx(&sm.full[r], bytes);
    bulk_g2s(sm.rows + (size_t)r * kRowBytes, table + ((size_t)__ldg(vp + ep) * VLN_V + r) * VLN_IMG, kRowBytes,
             &sm.full[r]);
    if (mask_bits)
      bulk_g2s(sm.mask + (size_t)r * kMaskBytes, mask_bits + ((size_t)ep * VLN_V + r) * kMaskBytes, kMaskBytes,
               &sm.full[r]);
  };
  // rows freed at release point k: rows 3k..3k+2 of every phase-2 group
  auto release_row = [](int k, int i) { return (i / 3) * kRowsPerGroup + 3 * k + (i % 3); };

  const float scale = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
  const uint32_t thr = drop_threshold(drop_p);
  uint64_t seed = 0, offset = 0;
  if (drop_p > 0.f && !mask_bits) {
    seed = rng[0];
    offset = rng[1] + call_off;
  }
  // stage the per-unit vectors (query, angle table row, saved attention) of episode `e` into buffer `buf`
  auto prefetch_unit = [&](int e, int buf) {
    if (e < B) {
      const float* vr = vec + (size_t)e * ld_vec;
      for (int c = tid; c < VLN_F / 4; c += kThreads) cp_async16(&sm.qbuf[buf][c * 4], vr + c * 4);
      const float* lr = loc4 + (size_t)__ldg(view + e) * (VLN_V * 4);
      if (tid < VLN_V) cp_async16(&sm.locbuf[buf][tid * 4], lr + tid * 4);
      if (mode == 1 && tid < VLN_V / 4) cp_async16(&sm.attbuf[buf][tid * 4], attn_io + (size_t)e * VLN_V + tid * 4);
    }
    cp_async_commit();
  };

  if ((int)blockIdx.x < B && warp == 0) {                  // first unit: all 36 rows at once
    for (int r = lane; r < VLN_V; r += 32) request_row(blockIdx.x, r);
  }
  prefetch_unit(blockIdx.x, 0);

  for (int ep = blockIdx.x; ep < B; ep += gridDim.x, ++it) {
    const int buf = it & 1;
    const uint32_t ph = (uint32_t)it & 1u;
    const int next_ep = ep + gridDim.x;
    cp_async_wait_all();
    __syncthreads();                                       // this unit's vectors are in place for every warp
    prefetch_unit(next_ep, buf ^ 1);                       // overlaps with this unit's work
    PSTAMP(2);

    // ------------------------------ phase 1: logits ------------------------------
    {
      // query slice in registers: q[j*8+e] = vec[j*256 + lane*8 + e]
      float q[64];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 a = *reinterpret_cast<const float4*>(&sm.qbuf[buf][j * 256 + lane * 8]);
        const float4 b = *reinterpret_cast<const float4*>(&sm.qbuf[buf][j * 256 + lane * 8 + 4]);
        q[j * 8 + 0] = a.x; q[j * 8 + 1] = a.y; q[j * 8 + 2] = a.z; q[j * 8 + 3] = a.w;
        q[j * 8 + 4] = b.x; q[j * 8 + 5] = b.y; q[j * 8 + 6] = b.z; q[j * 8 + 7] = b.w;
      }
      // angle group sums of the query: qa[k] = sum_i vec[2048 + 32k + i]
      float qa0, qa1, qa2, qa3;
      {
        const float4 a = *reinterpret_cast<const float4*>(&sm.qbuf[buf][VLN_IMG + lane * 4]);
        float s = a.x + a.y + a.z + a.w;
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        qa0 = __shfl_sync(0xffffffffu, s, 0);
        qa1 = __shfl_sync(0xffffffffu, s, 8);
        qa2 = __shfl_sync(0xffffffffu, s, 16);
        qa3 = __shfl_sync(0xffffffffu, s, 24);
      }
#pragma unroll 1
      for (int v = warp; v < VLN_V; v += kWarps) {
        mbar_wait(&sm.full[v], ph);
        uint4* rowp = reinterpret_cast<uint4*>(sm.rows + (size_t)v * kRowBytes);
        uint4 x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = rowp[j * 32 + lane];
        if (drop_p > 0.f) {
          if (mask_bits) {
            // pre-generated keep-bits: byte j of this lane's 8-byte group covers the 8 features of x[j]
            const uint2 mb = *reinterpret_cast<const uint2*>(sm.mask + (size_t)v * kMaskBytes + lane * 8);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t bits = ((j < 4 ? mb.x : mb.y) >> ((j & 3) * 8)) & 0xFFu;
              uint32_t* w = reinterpret_cast<uint32_t*>(&x[j]);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                w[k] &= (((bits >> (2 * k)) & 1u) * 0x0000FFFFu) | (((bits >> (2 * k + 1)) & 1u) * 0xFFFF0000u);
            }
          } else {
            const uint64_t e0 = (((uint64_t)ep * VLN_V + (uint64_t)v) * VLN_IMG) >> 3;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const Philox8 r = philox8(seed, offset, e0 + (uint64_t)(j * 32 + lane));
              uint32_t* w = reinterpret_cast<uint32_t*>(&x[j]);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                w[k] &= (philox_keep(r, 2 * k, thr) ? 0x0000FFFFu : 0u) | (philox_keep(r, 2 * k + 1, thr) ? 0xFFFF0000u : 0u);
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) rowp[j * 32 + lane] = x[j];     // phase 2 reads the masked row
        }
        float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;         // four chains: no 64-deep FMA dependency
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          d0 = fmaf(bf16lo(x[j].x), q[j * 8 + 0], d0); d1 = fmaf(bf16hi(x[j].x), q[j * 8 + 1], d1);
          d2 = fmaf(bf16lo(x[j].y), q[j * 8 + 2], d2); d3 = fmaf(bf16hi(x[j].y), q[j * 8 + 3], d3);
          d0 = fmaf(bf16lo(x[j].z), q[j * 8 + 4], d0); d1 = fmaf(bf16hi(x[j].z), q[j * 8 + 5], d1);
          d2 = fmaf(bf16lo(x[j].w), q[j * 8 + 6], d2); d3 = fmaf(bf16hi(x[j].w), q[j * 8 + 7], d3);
        }
        float dot = warp_sum((d0 + d1) + (d2 + d3)) * scale;
        const float4 lv = *reinterpret_cast<const float4*>(&sm.locbuf[buf][v * 4]);   // this view's angle values
        dot += lv.x * qa0 + lv.y * qa1 + lv.z * qa2 + lv.w * qa3;
        if (lane == 0) sm.logit[v] = dot;
      }
    }
    PSTAMP(3);
    __syncthreads();
    PSTAMP(4);

    // ------------------------------ softmax / backward coefficients ------------------------------
    if (warp == 0) {
      const float x0 = sm.logit[lane], x1 = lane < VLN_V - 32 ? sm.logit[32 + lane] : -INFINITY;
      float w0, w1;
      if (mode == 0) {
        const float m = warp_max(fmaxf(x0, x1));
        const float e0 = __expf(x0 - m), e1 = lane < VLN_V - 32 ? __expf(x1 - m) : 0.f;
        const float inv = 1.0f / warp_sum(e0 + e1);
        w0 = e0 * inv;
        w1 = e1 * inv;
        attn_io[(size_t)ep * VLN_V + lane] = w0;
        if (lane < VLN_V - 32) attn_io[(size_t)ep * VLN_V + 32 + lane] = w1;
      } else {
        const float a0 = sm.attbuf[buf][lane], a1 = lane < VLN_V - 32 ? sm.attbuf[buf][32 + lane] : 0.f;
        const float rbar = warp_sum(a0 * x0 + (lane < VLN_V - 32 ? a1 * x1 : 0.f));
        w0 = a0 * (x0 - rbar);
        w1 = lane < VLN_V - 32 ? a1 * (x1 - rbar) : 0.f;
      }
      sm.wv[lane] = w0;
      if (lane < VLN_V - 32) sm.wv[32 + lane] = w1;
      // the 128 angle dimensions of the output: 4 values, each repeated x32
      float g0 = w0 * sm.locbuf[buf][lane * 4 + 0], g1 = w0 * sm.locbuf[buf][lane * 4 + 1];
      float g2 = w0 * sm.locbuf[buf][lane * 4 + 2], g3 = w0 * sm.locbuf[buf][lane * 4 + 3];
      if (lane < VLN_V - 32) {
        g0 += w1 * sm.locbuf[buf][(32 + lane) * 4 + 0]; g1 += w1 * sm.locbuf[buf][(32 + lane) * 4 + 1];
        g2 += w1 * sm.locbuf[buf][(32 + lane) * 4 + 2]; g3 += w1 * sm.locbuf[buf][(32 + lane) * 4 + 3];
      }
      g0 = warp_sum(g0); g1 = warp_sum(g1); g2 = warp_sum(g2); g3 = warp_sum(g3);
      float* orow = out + (size_t)ep * ld_out + VLN_IMG;
      orow[lane] = g0; orow[32 + lane] = g1; orow[64 + lane] = g2; orow[96 + lane] = g3;
    }
    __syncthreads();
    PSTAMP(5);

    // ------------------------------ phase 2: weighted sum, column-parallel ------------------------------
    {
      const int g = tid >> 7, t = tid & 127;               // group g: rows 12g..12g+11; chunks t and t+128
      float acc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = 0.f;
#pragma unroll 1
      for (int k = 0; k < kRel; ++k) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int v = g * kRowsPerGroup + 3 * k + i;
          const float w = sm.wv[v];
          const uint4* rowp = reinterpret_cast<const uint4*>(sm.rows + (size_t)v * kRowBytes);
          const uint4 xa = rowp[t], xb = rowp[t + 128];
          acc[0] = fmaf(w, bf16lo(xa.x), acc[0]); acc[1] = fmaf(w, bf16hi(xa.x), acc[1]);
          acc[2] = fmaf(w, bf16lo(xa.y), acc[2]); acc[3] = fmaf(w, bf16hi(xa.y), acc[3]);
          acc[4] = fmaf(w, bf16lo(xa.z), acc[4]); acc[5] = fmaf(w, bf16hi(xa.z), acc[5]);
          acc[6] = fmaf(w, bf16lo(xa.w), acc[6]); acc[7] = fmaf(w, bf16hi(xa.w), acc[7]);
          acc[8] = fmaf(w, bf16lo(xb.x), acc[8]); acc[9] = fmaf(w, bf16hi(xb.x), acc[9]);
          acc[10] = fmaf(w, bf16lo(xb.y), acc[10]); acc[11] = fmaf(w, bf16hi(xb.y), acc[11]);
          acc[12] = fmaf(w, bf16lo(xb.z), acc[12]); acc[13] = fmaf(w, bf16hi(xb.z), acc[13]);
          acc[14] = fmaf(w, bf16lo(xb.w), acc[14]); acc[15] = fmaf(w, bf16hi(xb.w), acc[15]);
        }
        // release point k: rows 3k..3k+2 of every group are done; refill them with the next episode's rows
        __syncthreads();
        if (next_ep < B && warp == 0 && lane < 9) {
          fence_proxy_async();                             // order the generic-proxy accesses before the async writes
          request_row(next_ep, release_row(k, lane));
        }
      }
      PSTAMP(6);
      if (g > 0) {
        float* p = sm.part + ((size_t)(g - 1) * 128 + t) * 16;
#pragma unroll
        for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(p + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
      }
      __syncthreads();
      if (g == 0) {
#pragma unroll
        for (int gg = 0; gg < kGroups - 1; ++gg) {
          const float* p = sm.part + ((size_t)gg * 128 + t) * 16;
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 a = *reinterpret_cast<const float4*>(p + i);
            acc[i] += a.x; acc[i + 1] += a.y; acc[i + 2] += a.z; acc[i + 3] += a.w;
          }
        }
        float* orow = out + (size_t)ep * ld_out;
        reinterpret_cast<float4*>(orow + t * 8)[0] = make_float4(acc[0] * scale, acc[1] * scale, acc[2] * scale, acc[3] * scale);
        reinterpret_cast<float4*>(orow + t * 8)[1] = make_float4(acc[4] * scale, acc[5] * scale, acc[6] * scale, acc[7] * scale);
        reinterpret_cast<float4*>(orow + (t + 128) * 8)[0] = make_float4(acc[8] * scale, acc[9] * scale, acc[10] * scale, acc[11] * scale);
        reinterpret_cast<float4*>(orow + (t + 128) * 8)[1] = make_float4(acc[12] * scale, acc[13] * scale, acc[14] * scale, acc[15] * scale);
      }
    }
    PSTAMP(7);
    // sm.part / logit / wv are rewritten only after the next iteration's first __syncthreads
  }
}

}  // namespace

extern "C" int vln_pano_attn_ld(const vln_ctx* ctx, const int32_t* vp, const int32_t* view, const float* loc4,
                                const float* vec, int ld_vec, float* attn_io, const float* fwd_out, int ld_fwd,
                                float* out, int ld_out, int B, int mode, float drop_p, const uint64_t* rng,
                                uint64_t call_off, const uint8_t* mask_bits, int split, void* stream) {
  VLN_REQUIRE(ctx && vp && view && loc4 && vec && attn_io && out && B > 0, "bad arguments");
  VLN_REQUIRE(split == 1 || split == 2 || split == 4, "split must be 1, 2 or 4 (kept for ABI stability; unused since v3)");
  VLN_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (forward) or 1 (backward)");
  VLN_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "drop_p out of range");
  VLN_REQUIRE(drop_p == 0.f || rng || mask_bits, "dropout needs an rng state or pre-generated keep-bits");
  VLN_REQUIRE(!mask_bits || (drop_p > 0.f && ((uintptr_t)mask_bits & 15) == 0), "mask_bits: 16-byte aligned, with drop_p > 0");
  VLN_REQUIRE(ld_vec >= VLN_F && ld_out >= VLN_F, "row strides must be >= 2176");
  VLN_REQUIRE(ld_vec % 4 == 0 && ld_out % 4 == 0 && ((uintptr_t)vec & 15) == 0 && ((uintptr_t)out & 15) == 0,
              "rows must be 16-byte aligned");
  (void)fwd_out;
  (void)ld_fwd;
  static bool configured = false;
  if (!configured) {
    VLN_CHECK_CUDA(cudaFuncSetAttribute(pano_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
    configured = true;
  }
  const int grid = B < ctx->num_sms ? B : ctx->num_sms;
  VLN_CHECK_CUDA(vln_launch_chain(pano_attn_kernel, dim3(grid), dim3(kThreads), sizeof(Smem), (cudaStream_t)stream,
                                  ctx->table, vp, view, loc4, vec, attn_io, out, B, mode, drop_p, rng, call_off, ld_vec,
                                  ld_out, mask_bits, (int)(getenv("VLN_PANO_STAMPS") != nullptr)));
  return 0;
}

extern "C" int vln_debug_pano_stamps(unsigned long long* out_host /*[16]*/) {
  VLN_CHECK_CUDA(cudaDeviceSynchronize());
  VLN_CHECK_CUDA(cudaMemcpyFromSymbol(out_host, g_pano_stamps, sizeof(unsigned long long) * 16));
  return 0;
}

extern "C" int vln_pano_attn(const vln_ctx* ctx, const int32_t* vp, const int32_t* view, const float* loc4,
                             const float* vec, float* attn_io, const float* fwd_out, float* out, int B, int mode,
                             float drop_p, const uint64_t* rng, uint64_t call_off, int split, void* stream) {
  return vln_pano_attn_ld(ctx, vp, view, loc4, vec, VLN_F, attn_io, fwd_out, VLN_F, out, VLN_F, B, mode, drop_p, rng,
                          call_off, nullptr, split, stream);
}
