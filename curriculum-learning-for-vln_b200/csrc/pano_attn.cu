// Fused gather + feature dropout + soft-dot attention over the 36-view panorama
// (SoftDotAttention.forward units.py:107-118 on `_feature_variable` base.py:141-147 with
// EnvDropDecoder's env_drop policy.py:226-231), forward and backward in one kernel.
//
// Roofline class: HBM.  Algorithmic bytes: 36 x 2048 x 2 = 147 456 B per episode-step, read ONCE.
//
// Persistent, warp-specialised, single pass (flash-decoding style):
//   * work unit = (episode, part): `split` parts of 36/split views each, so that B=64 episodes
//     still spread over ~all SMs.  CTA c processes units c, c+grid, ...
//   * producer warp: one lane streams the unit's view rows (4 096 B each, contiguous in the table)
//     into a shared-memory ring with cp.async.bulk (TMA bulk copy, mbarrier complete_tx); it runs
//     up to kStages rows ahead of the consumers, across unit boundaries.
//   * 6 consumer warps: warp w takes rows w, w+6, ...; a row is read from shared memory exactly
//     once into registers (8 x 16 B per lane), the keep-mask is applied, the dot product with the
//     register-resident query is reduced with shuffles, the ring slot is released, and the row —
//     still in registers — is folded into a running (max, sum, weighted accumulator): online
//     softmax, so no second pass over the tile.  Backward needs no softmax state at all:
//         dq = sum_v a_v r_v x_v - (sum_v a_v r_v) * out_fwd,     r_v = x_v . d_out
//   * unit end: the 6 warps' partials are merged through shared memory; with split > 1 the CTA
//     partials go to a global scratch slab and the LAST part to arrive (atomic ticket) merges them —
//     no cluster, no second launch.
// The 128 angle dimensions are 4 distinct values per view (misc.py:286-293): they ride along as 4
// scalars per row (loc4) against 4 group sums of the query.
#include "common.cuh"

namespace {

constexpr int kConsumerWarps = 6;                        // 36, 18 rows split evenly; 7 warps -> 255 regs/thread
constexpr int kConsumers = kConsumerWarps * 32;
constexpr int kThreads = kConsumers + 32;                // + producer warp
constexpr int kRowBytes = VLN_IMG * 2;                   // 4 096
constexpr int kMaskBytes = VLN_IMG / 8;                  // 256: packed keep-bits of one row (vln_feature_mask_bits)
constexpr int kStageBytes = kRowBytes + kMaskBytes;
constexpr int kStages = 32;                              // 136 KB ring; stage s is fed by producer lane s
constexpr int kRed = VLN_IMG + 8;                        // acc[2048], accA[4], m, l (or dsum), pad
constexpr int kSlab = VLN_IMG + 8;                       // floats per (episode, part) scratch slab
constexpr int kChunks = VLN_IMG / 4;                     // float4 chunks of an output row

__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bar_consumers() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory"); }

struct Smem {
  uint8_t ring[kStages * kStageBytes];
  float red[kConsumerWarps * kRed];
  float qbuf[2][VLN_F];            // query / d_out row of this and the next unit (cp.async double buffer)
  float locbuf[2][VLN_V * 4];      // loc4[cur_view] rows
  float attbuf[2][40];             // saved attention (backward)
  float logit[40];
  float bcast[8];
  uint64_t full[kStages];
  uint64_t empty[kStages];
};

__global__ void __launch_bounds__(kThreads, 1)
pano_attn_kernel(const __nv_bfloat16* __restrict__ table, const int32_t* __restrict__ vp,
                 const int32_t* __restrict__ view, const float* __restrict__ loc4, const float* __restrict__ vec,
                 float* __restrict__ attn_io, const float* __restrict__ fwd_out, float* __restrict__ out,
                 float* __restrict__ scratch, unsigned int* __restrict__ tickets, int B, int S, int mode,
                 float drop_p, const uint64_t* __restrict__ rng, uint64_t call_off, int ld_vec, int ld_out,
                 int ld_fwd, const uint8_t* __restrict__ mask_bits) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int R = VLN_V / S;                                // rows per unit
  const int n_units = B * S;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], 1);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == kConsumerWarps) {
    // ------------------------------- producer -------------------------------
    // lane s owns ring stage s: it streams rows s, s+32, ... of this CTA's row sequence, so 32 bulk
    // copies are issued per warp iteration instead of one.
    // Lanes poll their own `empty` barrier without blocking, so a lane whose slot is still in
    // use never holds back the others (a blocking wait would reconverge the warp on the slowest slot).
    uint32_t n = lane;
    const __nv_bfloat16* src = nullptr;
    const uint8_t* msrc = nullptr;
    auto locate = [&]() -> bool {
      const uint32_t k = n / (uint32_t)R, i = n - k * (uint32_t)R;
      const int u = blockIdx.x + (int)k * (int)gridDim.x;
      if (u >= n_units) return false;
      const int ep = u / S, part = u - ep * S;
      src = table + ((size_t)__ldg(vp + ep) * VLN_V + (size_t)part * R + i) * VLN_IMG;
      if (mask_bits) msrc = mask_bits + ((size_t)ep * VLN_V + (size_t)part * R + i) * kMaskBytes;
      return true;
    };
    bool work = locate();
    while (__any_sync(0xffffffffu, work)) {
      if (work && mbar_test_wait(&sm.empty[lane], ((n / kStages) & 1u) ^ 1u)) {
        mbar_expect_tx(&sm.full[lane], mask_bits ? kStageBytes : kRowBytes);
        bulk_g2s(sm.ring + (size_t)lane * kStageBytes, src, kRowBytes, &sm.full[lane]);
        if (mask_bits) bulk_g2s(sm.ring + (size_t)lane * kStageBytes + kRowBytes, msrc, kMaskBytes, &sm.full[lane]);
        n += kStages;
        work = locate();
      }
    }
    return;
  }

  // --------------------------------- consumers ---------------------------------
  const float scale = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
  const uint32_t thr = drop_threshold(drop_p);
  uint64_t seed = 0, offset = 0;
  if (drop_p > 0.f && !mask_bits) {
    seed = rng[0];
    offset = rng[1] + call_off;
  }
  // stage the per-unit vectors (query, angle table row, saved attention) of unit `uu` into buffer `buf`
  auto prefetch_unit = [&](int uu, int buf) {
    if (uu < n_units) {
      const int e = uu / S;
      const float* vr = vec + (size_t)e * ld_vec;
      for (int c = tid; c < VLN_F / 4; c += kConsumers) cp_async16(&sm.qbuf[buf][c * 4], vr + c * 4);
      const float* lr = loc4 + (size_t)__ldg(view + e) * (VLN_V * 4);
      if (tid < VLN_V) cp_async16(&sm.locbuf[buf][tid * 4], lr + tid * 4);
      if (mode == 1 && tid < VLN_V / 4) cp_async16(&sm.attbuf[buf][tid * 4], attn_io + (size_t)e * VLN_V + tid * 4);
    }
    cp_async_commit();
  };
  prefetch_unit(blockIdx.x, 0);
  uint32_t n_base = 0;
  int it = 0;
  for (int u = blockIdx.x; u < n_units; u += gridDim.x, n_base += R, ++it) {
    const int ep = u / S, part = u - ep * S;
    const int buf = it & 1;
    cp_async_wait_all();
    bar_consumers();                                       // this unit's vectors are in place for every warp
    prefetch_unit(u + gridDim.x, buf ^ 1);                 // overlaps with this unit's rows
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

    float acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) acc[i] = 0.f;
    float aA0 = 0.f, aA1 = 0.f, aA2 = 0.f, aA3 = 0.f;
    float m_run = -INFINITY, l_run = 0.f;                  // forward: running max / sum; backward: l_run = sum_v w_v

    for (int i = warp; i < R; i += kConsumerWarps) {
      const uint32_t n = n_base + (uint32_t)i;
      const uint32_t s = n % kStages, ph = (n / kStages) & 1u;
      const int v = part * R + i;
      mbar_wait(&sm.full[s], ph);
      const uint4* rowp = reinterpret_cast<const uint4*>(sm.ring + (size_t)s * kStageBytes);
      uint4 x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = rowp[j * 32 + lane];
      if (mask_bits) {
        // pre-generated keep-bits: byte j of this lane's 8-byte group covers the 8 features of x[j]
        const uint2 mb = *reinterpret_cast<const uint2*>(sm.ring + (size_t)s * kStageBytes + kRowBytes + lane * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t bits = ((j < 4 ? mb.x : mb.y) >> ((j & 3) * 8)) & 0xFFu;
          uint32_t* w = reinterpret_cast<uint32_t*>(&x[j]);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            w[k] &= (((bits >> (2 * k)) & 1u) * 0x0000FFFFu) | (((bits >> (2 * k + 1)) & 1u) * 0xFFFF0000u);
        }
      } else if (drop_p > 0.f) {
        const uint64_t e0 = (((uint64_t)ep * VLN_V + (uint64_t)v) * VLN_IMG) >> 3;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const Philox8 r = philox8(seed, offset, e0 + (uint64_t)(j * 32 + lane));
          uint32_t* w = reinterpret_cast<uint32_t*>(&x[j]);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t mk = (philox_keep(r, 2 * k, thr) ? 0x0000FFFFu : 0u) |
                                (philox_keep(r, 2 * k + 1, thr) ? 0xFFFF0000u : 0u);
            w[k] &= mk;
          }
        }
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
      if (lane == 0) mbar_arrive(&sm.empty[s]);            // every lane's loads fed `dot`: the slot is free
      const float4 lv = *reinterpret_cast<const float4*>(&sm.locbuf[buf][v * 4]);   // this view's angle values
      const float la = lv.x, lb = lv.y, lc = lv.z, ld = lv.w;
      dot += la * qa0 + lb * qa1 + lc * qa2 + ld * qa3;
      float w;
      if (mode == 0) {
        if (lane == 0) sm.logit[i] = dot;
        if (dot > m_run) {                                 // warp-uniform
          const float c = __expf(m_run - dot);             // exp(-inf) = 0 on the first row
          l_run *= c;
          aA0 *= c; aA1 *= c; aA2 *= c; aA3 *= c;
#pragma unroll
          for (int k = 0; k < 64; ++k) acc[k] *= c;
          m_run = dot;
        }
        w = __expf(dot - m_run);
      } else {
        w = sm.attbuf[buf][v] * dot;
      }
      l_run += w;
      aA0 += w * la; aA1 += w * lb; aA2 += w * lc; aA3 += w * ld;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[j * 8 + 0] += w * bf16lo(x[j].x); acc[j * 8 + 1] += w * bf16hi(x[j].x);
        acc[j * 8 + 2] += w * bf16lo(x[j].y); acc[j * 8 + 3] += w * bf16hi(x[j].y);
        acc[j * 8 + 4] += w * bf16lo(x[j].z); acc[j * 8 + 5] += w * bf16hi(x[j].z);
        acc[j * 8 + 6] += w * bf16lo(x[j].w); acc[j * 8 + 7] += w * bf16hi(x[j].w);
      }
    }

    // ---- merge the 8 warps ----
    float* rw = sm.red + (size_t)warp * kRed;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      reinterpret_cast<float4*>(rw + j * 256 + lane * 8)[0] = make_float4(acc[j * 8], acc[j * 8 + 1], acc[j * 8 + 2], acc[j * 8 + 3]);
      reinterpret_cast<float4*>(rw + j * 256 + lane * 8)[1] = make_float4(acc[j * 8 + 4], acc[j * 8 + 5], acc[j * 8 + 6], acc[j * 8 + 7]);
    }
    if (lane == 0) {
      rw[VLN_IMG + 0] = aA0; rw[VLN_IMG + 1] = aA1; rw[VLN_IMG + 2] = aA2; rw[VLN_IMG + 3] = aA3;
      rw[VLN_IMG + 4] = m_run; rw[VLN_IMG + 5] = l_run;
    }
    bar_consumers();
    // ---- merge: float4 chunk c of the output row is handled by consumer thread c % kConsumers ----
    float wsc[kConsumerWarps];
    float M = -INFINITY, L = 0.f;
    if (mode == 0) {
#pragma unroll
      for (int w = 0; w < kConsumerWarps; ++w) M = fmaxf(M, sm.red[w * kRed + VLN_IMG + 4]);
#pragma unroll
      for (int w = 0; w < kConsumerWarps; ++w) {
        const float mw = sm.red[w * kRed + VLN_IMG + 4];
        wsc[w] = mw == -INFINITY ? 0.f : __expf(mw - M);
        L += wsc[w] * sm.red[w * kRed + VLN_IMG + 5];
      }
    } else {
#pragma unroll
      for (int w = 0; w < kConsumerWarps; ++w) {
        wsc[w] = 1.f;
        L += sm.red[w * kRed + VLN_IMG + 5];
      }
    }
    float oA = 0.f;                                        // threads 0..3: angle group tid
    if (tid < 4) {
#pragma unroll
      for (int w = 0; w < kConsumerWarps; ++w) oA += wsc[w] * sm.red[w * kRed + VLN_IMG + tid];
    }
    float* orow = out + (size_t)ep * ld_out;
    const float* frow = mode == 1 ? fwd_out + (size_t)ep * ld_fwd : nullptr;
    bool finalize = true;
    if (S == 1) {
      const float f1 = mode == 0 ? scale / L : scale;
      for (int c = tid; c < kChunks; c += kConsumers) {
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int w = 0; w < kConsumerWarps; ++w) {
          const float4 a = reinterpret_cast<const float4*>(sm.red + w * kRed)[c];
          o.x += wsc[w] * a.x; o.y += wsc[w] * a.y; o.z += wsc[w] * a.z; o.w += wsc[w] * a.w;
        }
        if (mode == 0) {
          o.x *= f1; o.y *= f1; o.z *= f1; o.w *= f1;
        } else {
          const float4 f = __ldg(reinterpret_cast<const float4*>(frow) + c);
          o.x = o.x * f1 - L * f.x; o.y = o.y * f1 - L * f.y; o.z = o.z * f1 - L * f.z; o.w = o.w * f1 - L * f.w;
        }
        reinterpret_cast<float4*>(orow)[c] = o;
      }
      if (mode == 0)
        for (int i = tid; i < VLN_V; i += kConsumers) attn_io[(size_t)ep * VLN_V + i] = __expf(sm.logit[i] - M) / L;
    } else {
      // publish this part's partial, take a ticket; the last part to arrive merges all S partials
      float* slab = scratch + ((size_t)ep * S + part) * kSlab;
      for (int c = tid; c < kChunks; c += kConsumers) {
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int w = 0; w < kConsumerWarps; ++w) {
          const float4 a = reinterpret_cast<const float4*>(sm.red + w * kRed)[c];
          o.x += wsc[w] * a.x; o.y += wsc[w] * a.y; o.z += wsc[w] * a.z; o.w += wsc[w] * a.w;
        }
        reinterpret_cast<float4*>(slab)[c] = o;
      }
      if (tid < 4) slab[VLN_IMG + tid] = oA;
      if (tid == 0) {
        slab[VLN_IMG + 4] = M;
        slab[VLN_IMG + 5] = L;
      }
      if (mode == 0)                                       // un-normalised logits; the merger turns them into attention
        for (int i = tid; i < R; i += kConsumers) attn_io[(size_t)ep * VLN_V + part * R + i] = sm.logit[i];
      __threadfence();
      bar_consumers();
      if (tid == 0) {
        const unsigned int t = atomicAdd(tickets + ep, 1u);
        const bool last = t == (unsigned)S - 1u;
        if (last) tickets[ep] = 0u;                          // nobody else touches it any more in this launch
        sm.bcast[0] = last ? 1.f : 0.f;
      }
      bar_consumers();
      finalize = sm.bcast[0] != 0.f;
      if (finalize) {
        __threadfence();
        const float* eb = scratch + (size_t)ep * S * kSlab;
        float psc[4] = {0.f, 0.f, 0.f, 0.f};
        M = -INFINITY;
        L = 0.f;
        if (mode == 0) {
#pragma unroll
          for (int p = 0; p < 4; ++p)
            if (p < S) M = fmaxf(M, __ldcg(eb + p * kSlab + VLN_IMG + 4));
#pragma unroll
          for (int p = 0; p < 4; ++p)
            if (p < S) {
              psc[p] = __expf(__ldcg(eb + p * kSlab + VLN_IMG + 4) - M);
              L += psc[p] * __ldcg(eb + p * kSlab + VLN_IMG + 5);
            }
        } else {
#pragma unroll
          for (int p = 0; p < 4; ++p)
            if (p < S) {
              psc[p] = 1.f;
              L += __ldcg(eb + p * kSlab + VLN_IMG + 5);
            }
        }
        oA = 0.f;
        if (tid < 4) {
#pragma unroll
          for (int p = 0; p < 4; ++p)
            if (p < S) oA += psc[p] * __ldcg(eb + p * kSlab + VLN_IMG + tid);
        }
        const float f1 = mode == 0 ? scale / L : scale;
        for (int c = tid; c < kChunks; c += kConsumers) {
          float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int p = 0; p < 4; ++p)
            if (p < S) {
              const float4 a = __ldcg(reinterpret_cast<const float4*>(eb + p * kSlab) + c);
              o.x += psc[p] * a.x; o.y += psc[p] * a.y; o.z += psc[p] * a.z; o.w += psc[p] * a.w;
            }
          if (mode == 0) {
            o.x *= f1; o.y *= f1; o.z *= f1; o.w *= f1;
          } else {
            const float4 f = __ldg(reinterpret_cast<const float4*>(frow) + c);
            o.x = o.x * f1 - L * f.x; o.y = o.y * f1 - L * f.y; o.z = o.z * f1 - L * f.z; o.w = o.w * f1 - L * f.w;
          }
          reinterpret_cast<float4*>(orow)[c] = o;
        }
        if (mode == 0)
          for (int i = tid; i < VLN_V; i += kConsumers) {
            float* ap = attn_io + (size_t)ep * VLN_V + i;
            *ap = __expf(__ldcg(ap) - M) / L;
          }
      }
    }
    if (finalize) {                                        // the 128 angle dimensions: 4 values, each repeated x32
      if (tid < 4) sm.bcast[4 + tid] = mode == 0 ? oA / L : oA;
      bar_consumers();
      if (tid < VLN_ANG) {
        const float a = sm.bcast[4 + (tid >> 5)];
        orow[VLN_IMG + tid] = mode == 0 ? a : a - L * frow[VLN_IMG + tid];
      }
    }
    bar_consumers();                                       // red / logit / bcast are reused by the next unit
  }
}

}  // namespace

extern "C" int vln_pano_attn_ld(const vln_ctx* ctx, const int32_t* vp, const int32_t* view, const float* loc4,
                                const float* vec, int ld_vec, float* attn_io, const float* fwd_out, int ld_fwd,
                                float* out, int ld_out, int B, int mode, float drop_p, const uint64_t* rng,
                                uint64_t call_off, const uint8_t* mask_bits, int split, void* stream) {
  VLN_REQUIRE(ctx && vp && view && loc4 && vec && attn_io && out && B > 0, "bad arguments");
  VLN_REQUIRE(split == 1 || split == 2 || split == 4, "split must be 1, 2 or 4");
  VLN_REQUIRE(mode == 0 || (mode == 1 && fwd_out), "mode must be 0 (forward) or 1 (backward, needs fwd_out)");
  VLN_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "drop_p out of range");
  VLN_REQUIRE(drop_p == 0.f || rng || mask_bits, "dropout needs an rng state or pre-generated keep-bits");
  VLN_REQUIRE(!mask_bits || (drop_p > 0.f && ((uintptr_t)mask_bits & 15) == 0), "mask_bits: 16-byte aligned, with drop_p > 0");
  VLN_REQUIRE(split == 1 || B <= VLN_SPLIT_MAX_B, "split > 1 supports at most VLN_SPLIT_MAX_B episodes per call");
  VLN_REQUIRE(ld_vec >= VLN_F && ld_out >= VLN_F && (mode == 0 || ld_fwd >= VLN_F), "row strides must be >= 2176");
  VLN_REQUIRE(ld_vec % 4 == 0 && ld_out % 4 == 0 && ld_fwd % 4 == 0 && ((uintptr_t)vec & 15) == 0 &&
                  ((uintptr_t)out & 15) == 0 && ((uintptr_t)fwd_out & 15) == 0,
              "rows must be 16-byte aligned");
  static bool configured = false;
  if (!configured) {
    VLN_CHECK_CUDA(cudaFuncSetAttribute(pano_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
    configured = true;
  }
  const int units = B * split;
  const int grid = units < ctx->num_sms ? units : ctx->num_sms;
  pano_attn_kernel<<<grid, kThreads, sizeof(Smem), (cudaStream_t)stream>>>(
      ctx->table, vp, view, loc4, vec, attn_io, fwd_out, out, ctx->scratch, ctx->tickets, B, split, mode, drop_p, rng,
      call_off, ld_vec, ld_out, ld_fwd, mask_bits);
  VLN_LAUNCH_OK();
  return 0;
}

extern "C" int vln_pano_attn(const vln_ctx* ctx, const int32_t* vp, const int32_t* view, const float* loc4,
                             const float* vec, float* attn_io, const float* fwd_out, float* out, int B, int mode,
                             float drop_p, const uint64_t* rng, uint64_t call_off, int split, void* stream) {
  return vln_pano_attn_ld(ctx, vp, view, loc4, vec, VLN_F, attn_io, fwd_out, VLN_F, out, VLN_F, B, mode, drop_p, rng,
                          call_off, nullptr, split, stream);
}
