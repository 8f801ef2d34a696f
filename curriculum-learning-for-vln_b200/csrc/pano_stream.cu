// Streaming variant of the fused gather + feature dropout + 36-view soft-dot attention
// (SoftDotAttention.forward units.py:107-118 over `_feature_variable` base.py:141-147 with
// EnvDropDecoder's env_drop policy.py:226-231) for launches with MANY episodes (config 5 of
// BASELINE.json, B = 148 ... 2048): the bandwidth-bound regime.
//
// Roofline class: HBM.  Algorithmic bytes: 36 x 2048 x 2 = 147 456 B per episode-step, read ONCE
// (+ 9 216 B of packed keep-bits when the feature dropout is on).
//
// The low-latency kernel of pano_attn.cu keeps a whole (half-)panorama in shared memory and walks it
// in two barrier-separated phases; with one unit per SM in flight those phases serialise and it tops
// out near 0.3-0.4 of the HBM peak.  Here nothing is resident:
//   * CTA = 4 warps = one episode at a time, 3 CTAs per SM (independent episodes hide each other's
//     barriers and reductions); warp w owns feature columns [512 w, 512 w + 512), 16 per lane;
//   * the 36 rows arrive as 9 chunks of 4 rows (16 KB + 1 KB of keep-bits: two bulk-async copies on one
//     mbarrier) through a 4-slot ring; a chunk goes shared -> registers at once and its slot is re-armed
//     with the chunk four ahead right after the CTA barrier, so ~4 x 17 KB per CTA (200 KB per SM) are
//     always in flight and every table byte crosses shared memory exactly once;
//   * single pass, online softmax: per chunk the 4 warps publish 4 partial dot products each, meet at
//     one CTA barrier, and every warp rescales its running (max, sum, acc[16 per lane]) identically;
//     no 2048-wide merge exists because the column split gives each warp its own output columns;
//   * backward is one pass too: dq = sum_v a_v r_v x~_v - (sum_v a_v r_v) sum_v a_v x~_v, both sums
//     accumulated side by side;
//   * keep-bits expand to bf16-pair masks with PRMT's sign-replicate mode (one shift serves two
//     words), products run on the packed FFMA2 pipe (fma.rn.f32x2).
#include <cstdlib>

#include "common.cuh"

namespace {

constexpr int kR = 4;                       // rows per chunk
constexpr int kChunks = VLN_V / kR;         // 9
constexpr int kRowB = VLN_IMG * 2;          // 4 096
constexpr int kMaskB = VLN_IMG / 8;         // 256

template <int kS>
struct SmemS {
  uint8_t rows[kS][kR * kRowB];             // 48 KB
  uint8_t mask[kS][kR * kMaskB];            // 3 KB
  float part[2][kR][4];                     // partial dots [chunk parity][row][warp]
  float slog[VLN_V + 4];                    // logits of the episode (forward: attention output)
  uint64_t full[kS];
};

__device__ __forceinline__ void bulk_g2s_s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
// d += a * b on both halves (packed fp32 pipe)
__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
  uint64_t dd = *reinterpret_cast<uint64_t*>(&d);
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(dd)
      : "l"(*reinterpret_cast<const uint64_t*>(&a)), "l"(*reinterpret_cast<const uint64_t*>(&b)));
  d = *reinterpret_cast<float2*>(&dd);
}
__device__ __forceinline__ void fmul2(float2& d, const float2 b) {
  uint64_t dd = *reinterpret_cast<uint64_t*>(&d);
  asm("mul.rn.f32x2 %0, %0, %1;" : "+l"(dd) : "l"(*reinterpret_cast<const uint64_t*>(&b)));
  d = *reinterpret_cast<float2*>(&dd);
}
__device__ __forceinline__ float2 bf2(uint32_t w) {   // bf16 pair -> two fp32
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xFFFF0000u));
}

// MODE 0 forward, 1 backward; MASK: pre-generated keep-bits streamed next to the rows
// kS ring slots, kCtas resident CTAs per SM: (4, 3) runs without spills at <= 168 registers; (3, 4) has to fit 128
// registers, spills, and measured 0.40-0.52 of peak where (4, 3) reaches 0.71-0.78
template <int MODE, bool MASK, int kS, int kCtas>
__global__ void __launch_bounds__(128, kCtas)
pano_stream_kernel(const __nv_bfloat16* __restrict__ table, const int32_t* __restrict__ vp,
                   const int32_t* __restrict__ view, const float* __restrict__ loc4, const float* __restrict__ vec,
                   float* __restrict__ attn_io, float* __restrict__ out, int B, float scale, int ld_vec, int ld_out,
                   const uint8_t* __restrict__ mask_bits, const __grid_constant__ ChainLink link) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  SmemS<kS>& sm = *reinterpret_cast<SmemS<kS>*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_trigger();
  if (tid < kS) mbar_init(&sm.full[tid], 1);
  fence_mbar_init();
  __syncthreads();
  chain_wait_cta(link);                                    // viewpoints, query and keep-bits come from predecessors

  const int stride = gridDim.x;
  int e = blockIdx.x;
  if (e >= B) {
    chain_signal_cta(link);
    return;
  }
  // one thread feeds the ring: chunk c of episode `ep` (viewpoint g) into `slot`
  auto issue = [&](int ep, int g, int c, int slot) {
    uint64_t* bar = &sm.full[slot];
    mbar_expect_tx(bar, MASK ? kR * (kRowB + kMaskB) : kR * kRowB);
    bulk_g2s_s(sm.rows[slot], table + ((size_t)g * VLN_V + (size_t)c * kR) * VLN_IMG, kR * kRowB, bar);
    if (MASK) bulk_g2s_s(sm.mask[slot], mask_bits + ((size_t)ep * VLN_V + (size_t)c * kR) * kMaskB, kR * kMaskB, bar);
  };
  int g_cur = __ldg(vp + e), g_next = 0;
  if (tid == 96) {
#pragma unroll
    for (int c = 0; c < kS; ++c) issue(e, g_cur, c, c);
  }
  int slot = 0, pb = 0;                                    // ring slot; parity of the running chunk count (part buffer)
  uint32_t par = 0;
  const int colw = warp * 512 + lane * 8;                  // first of this lane's columns (chunk j adds 256 j)

  for (; e < B; e += stride) {
    const int e_next = e + stride;
    if (e_next < B) g_next = __ldg(vp + e_next);
    // ---- this episode's vector: 16 columns per lane + (warp 0) the 4 angle group sums -------------------
    float2 q[8];
    {
      const float* vr = vec + (size_t)e * ld_vec + colw;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(vr + j * 256));
        const float4 b = __ldg(reinterpret_cast<const float4*>(vr + j * 256 + 4));
        q[j * 4 + 0] = make_float2(a.x, a.y); q[j * 4 + 1] = make_float2(a.z, a.w);
        q[j * 4 + 2] = make_float2(b.x, b.y); q[j * 4 + 3] = make_float2(b.z, b.w);
      }
    }
    float qa0 = 0.f, qa1 = 0.f, qa2 = 0.f, qa3 = 0.f;
    const float* locr = loc4;
    if (warp == 0) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(vec + (size_t)e * ld_vec + VLN_IMG + lane * 4));
      float s = (a.x + a.y) + (a.z + a.w);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      qa0 = __shfl_sync(0xffffffffu, s, 0);
      qa1 = __shfl_sync(0xffffffffu, s, 8);
      qa2 = __shfl_sync(0xffffffffu, s, 16);
      qa3 = __shfl_sync(0xffffffffu, s, 24);
      locr = loc4 + (size_t)__ldg(view + e) * (VLN_V * 4);
    }
    float m_run = -INFINITY, l_run = 0.f;                  // forward: running max / sum; backward: l_run = sum a_v r_v
    float2 acc[8], acc2[8];
    float an0 = 0.f, an1 = 0.f, an2 = 0.f, an3 = 0.f;      // angle part (warp 0)
    float bn0 = 0.f, bn1 = 0.f, bn2 = 0.f, bn3 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[i] = make_float2(0.f, 0.f);
      acc2[i] = make_float2(0.f, 0.f);
    }

#pragma unroll 1
    for (int c = 0; c < kChunks; ++c) {
      float av[kR];                                        // backward: saved attention of the chunk's rows
      if (MODE == 1) {
#pragma unroll
        for (int r = 0; r < kR; ++r) av[r] = __ldg(attn_io + (size_t)e * VLN_V + c * kR + r);
      }
      float4 lv[kR];
      if (warp == 0) {
#pragma unroll
        for (int r = 0; r < kR; ++r) lv[r] = __ldg(reinterpret_cast<const float4*>(locr + (c * kR + r) * 4));
      }
      mbar_wait(&sm.full[slot], par);
      // ---- shared -> registers, keep-mask applied on the packed words ----------------------------------
      uint32_t xw[kR][8];
#pragma unroll
      for (int r = 0; r < kR; ++r) {
        const uint8_t* rowp = sm.rows[slot] + r * kRowB + warp * 1024 + lane * 16;
        const uint4 a = *reinterpret_cast<const uint4*>(rowp), b = *reinterpret_cast<const uint4*>(rowp + 512);
        xw[r][0] = a.x; xw[r][1] = a.y; xw[r][2] = a.z; xw[r][3] = a.w;
        xw[r][4] = b.x; xw[r][5] = b.y; xw[r][6] = b.z; xw[r][7] = b.w;
      }
      if (MASK) {
#pragma unroll
        for (int r = 0; r < kR; ++r) {
          // bytes 0/1 = keep-bits of the lane's two 8-column groups (vln_feature_mask_bits layout)
          const uint32_t mb = *reinterpret_cast<const uint16_t*>(sm.mask[slot] + r * kMaskB + (warp >> 1) * 128 +
                                                                 lane * 4 + (warp & 1) * 2);
          // t: b | b << 9 for each byte, second byte's copy in the upper half.  After << (6 - 2k) the sign bits
          // of bytes 1, 0 are keep-bits 2k, 2k+1 of group 0 and those of bytes 3, 2 the same of group 1.
          const uint32_t t = (((mb & 0xFFu) * 0x201u) & 0xFFFFu) | ((mb >> 8) * 0x02010000u);   // (bit 16 belongs to group 1)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t ts = t << (6 - 2 * k);
            xw[r][k] &= prmt(ts, 0u, 0x8899u);
            xw[r][4 + k] &= prmt(ts, 0u, 0xAABBu);
          }
        }
      }
      // ---- partial dot products over my 16 columns, 4 rows reduced across the warp together -------------
      float p[kR];
#pragma unroll
      for (int r = 0; r < kR; ++r) {
        float2 d0 = make_float2(0.f, 0.f), d1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          ffma2(d0, bf2(xw[r][i]), q[i]);
          ffma2(d1, bf2(xw[r][i + 1]), q[i + 1]);
        }
        p[r] = (d0.x + d0.y) + (d1.x + d1.y);
      }
      {
        // lanes 0-15 keep rows 0,1 / lanes 16-31 rows 2,3; then bit 3 picks one of the two; 3 more butterflies
        const bool hi16 = lane & 16, hi8 = lane & 8;
        float k0 = hi16 ? p[2] : p[0], s0 = hi16 ? p[0] : p[2];
        float k1 = hi16 ? p[3] : p[1], s1 = hi16 ? p[1] : p[3];
        k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
        k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
        float k = hi8 ? k1 : k0, s = hi8 ? k0 : k1;
        k += __shfl_xor_sync(0xffffffffu, s, 8);
        k += __shfl_xor_sync(0xffffffffu, k, 4);
        k += __shfl_xor_sync(0xffffffffu, k, 2);
        k += __shfl_xor_sync(0xffffffffu, k, 1);
        if ((lane & 7) == 0) {
          const int r = (lane >> 4) * 2 + ((lane >> 3) & 1);
          float val = k * scale;
          if (warp == 0) {                                 // + angle part of the logit (4 values x 4 group sums)
            const float4 l4 = r == 0 ? lv[0] : (r == 1 ? lv[1] : (r == 2 ? lv[2] : lv[3]));
            val += l4.x * qa0 + l4.y * qa1 + l4.z * qa2 + l4.w * qa3;
          }
          sm.part[pb][r][warp] = val;
        }
      }
      __syncthreads();                                     // partials published; every warp has left the slot
      // Re-arm the slot with the chunk kS ahead.  Warp 3 does it (warp 0 carries the angle part, warp 1 the logits);
      // as in any TMA pipeline the consumers' reads are ordered before the refill by the barrier alone.
      if (tid == 96) {
        const int cn = c + kS;
        if (cn < kChunks) issue(e, g_cur, cn, slot);
        else if (e_next < B) issue(e_next, g_next, cn - kChunks, slot);
      }
      float sv[kR];
#pragma unroll
      for (int r = 0; r < kR; ++r) {
        const float4 pp = *reinterpret_cast<const float4*>(sm.part[pb][r]);
        sv[r] = (pp.x + pp.y) + (pp.z + pp.w);
      }
      if (MODE == 0) {
        const float m_new = fmaxf(fmaxf(m_run, fmaxf(sv[0], sv[1])), fmaxf(sv[2], sv[3]));
        const float corr = __expf(m_run - m_new);
        m_run = m_new;
        if (warp == 1 && lane < kR) sm.slog[c * kR + lane] = lane == 0 ? sv[0] : (lane == 1 ? sv[1] : (lane == 2 ? sv[2] : sv[3]));
        const float2 c2 = make_float2(corr, corr);
#pragma unroll
        for (int i = 0; i < 8; ++i) fmul2(acc[i], c2);
        l_run *= corr;
        if (warp == 0) { an0 *= corr; an1 *= corr; an2 *= corr; an3 *= corr; }
#pragma unroll
        for (int r = 0; r < kR; ++r) {
          const float w = __expf(sv[r] - m_new);
          l_run += w;
          const float2 w2 = make_float2(w, w);
#pragma unroll
          for (int i = 0; i < 8; ++i) ffma2(acc[i], bf2(xw[r][i]), w2);
          if (warp == 0) {
            an0 = fmaf(w, lv[r].x, an0); an1 = fmaf(w, lv[r].y, an1);
            an2 = fmaf(w, lv[r].z, an2); an3 = fmaf(w, lv[r].w, an3);
          }
        }
      } else {
#pragma unroll
        for (int r = 0; r < kR; ++r) {
          const float a = av[r], t = a * sv[r];
          l_run += t;
          const float2 a2 = make_float2(a, a), t2 = make_float2(t, t);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 x = bf2(xw[r][i]);
            ffma2(acc[i], x, t2);
            ffma2(acc2[i], x, a2);
          }
          if (warp == 0) {
            an0 = fmaf(t, lv[r].x, an0); an1 = fmaf(t, lv[r].y, an1);
            an2 = fmaf(t, lv[r].z, an2); an3 = fmaf(t, lv[r].w, an3);
            bn0 = fmaf(a, lv[r].x, bn0); bn1 = fmaf(a, lv[r].y, bn1);
            bn2 = fmaf(a, lv[r].z, bn2); bn3 = fmaf(a, lv[r].w, bn3);
          }
        }
      }
      pb ^= 1;                                             // (9 chunks per episode: c & 1 would repeat across episodes)
      if (++slot == kS) {
        slot = 0;
        par ^= 1u;
      }
    }

    // ---- episode epilogue --------------------------------------------------------------------------------
    float* orow = out + (size_t)e * ld_out + colw;
    if (MODE == 0) {
      const float inv = 1.0f / l_run, f = inv * scale;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        reinterpret_cast<float4*>(orow + j * 256)[0] =
            make_float4(acc[j * 4].x * f, acc[j * 4].y * f, acc[j * 4 + 1].x * f, acc[j * 4 + 1].y * f);
        reinterpret_cast<float4*>(orow + j * 256)[1] =
            make_float4(acc[j * 4 + 2].x * f, acc[j * 4 + 2].y * f, acc[j * 4 + 3].x * f, acc[j * 4 + 3].y * f);
      }
      if (warp == 0) {
        float* ar = out + (size_t)e * ld_out + VLN_IMG + lane;
        ar[0] = an0 * inv; ar[32] = an1 * inv; ar[64] = an2 * inv; ar[96] = an3 * inv;
      }
      if (warp == 1) {                                     // only warp 1 touches slog: program order suffices
        __syncwarp();
        attn_io[(size_t)e * VLN_V + lane] = __expf(sm.slog[lane] - m_run) * inv;
        if (lane < VLN_V - 32) attn_io[(size_t)e * VLN_V + 32 + lane] = __expf(sm.slog[32 + lane] - m_run) * inv;
        __syncwarp();
      }
    } else {
      const float rbar = l_run;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          o[2 * i] = (acc[j * 4 + i].x - rbar * acc2[j * 4 + i].x) * scale;
          o[2 * i + 1] = (acc[j * 4 + i].y - rbar * acc2[j * 4 + i].y) * scale;
        }
        reinterpret_cast<float4*>(orow + j * 256)[0] = make_float4(o[0], o[1], o[2], o[3]);
        reinterpret_cast<float4*>(orow + j * 256)[1] = make_float4(o[4], o[5], o[6], o[7]);
      }
      if (warp == 0) {
        float* ar = out + (size_t)e * ld_out + VLN_IMG + lane;
        ar[0] = an0 - rbar * bn0; ar[32] = an1 - rbar * bn1; ar[64] = an2 - rbar * bn2; ar[96] = an3 - rbar * bn3;
      }
    }
    g_cur = g_next;
  }
  chain_signal_cta(link);
}

template <int MODE, bool MASK, int kS, int kCtas>
cudaError_t launch_stream_cfg(const vln_ctx* ctx, const int32_t* vp, const int32_t* view, const float* loc4, const float* vec,
                          int ld_vec, float* attn_io, float* out, int ld_out, int B, float scale,
                          const uint8_t* mask_bits, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(pano_stream_kernel<MODE, MASK, kS, kCtas>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(SmemS<kS>));
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int max_ctas = ctx->num_sms * kCtas;
  const int grid = B < max_ctas ? B : max_ctas;
  const ChainLink link = vln_chain_link(stream, (unsigned int)grid);
  return vln_launch_linked(pano_stream_kernel<MODE, MASK, kS, kCtas>, dim3(grid), dim3(128), sizeof(SmemS<kS>), stream, ctx->table, vp, view,
                           loc4, vec, attn_io, out, B, scale, ld_vec, ld_out, mask_bits, link);
}

template <int MODE, bool MASK>
cudaError_t launch_stream(const vln_ctx* ctx, const int32_t* vp, const int32_t* view, const float* loc4, const float* vec,
                          int ld_vec, float* attn_io, float* out, int ld_out, int B, float scale,
                          const uint8_t* mask_bits, cudaStream_t stream) {
  return launch_stream_cfg<MODE, MASK, 4, 3>(ctx, vp, view, loc4, vec, ld_vec, attn_io, out, ld_out, B, scale, mask_bits, stream);
}

}  // namespace

// Streaming variant; inline-Philox dropout (drop_p > 0 without keep-bits) is not served here.
cudaError_t vln_pano_stream_launch(const vln_ctx* ctx, const int32_t* vp, const int32_t* view, const float* loc4,
                                   const float* vec, int ld_vec, float* attn_io, float* out, int ld_out, int B, int mode,
                                   float drop_p, const uint8_t* mask_bits, cudaStream_t stream) {
  const float scale = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
  if (mode == 0)
    return mask_bits ? launch_stream<0, true>(ctx, vp, view, loc4, vec, ld_vec, attn_io, out, ld_out, B, scale, mask_bits, stream)
                     : launch_stream<0, false>(ctx, vp, view, loc4, vec, ld_vec, attn_io, out, ld_out, B, scale, nullptr, stream);
  return mask_bits ? launch_stream<1, true>(ctx, vp, view, loc4, vec, ld_vec, attn_io, out, ld_out, B, scale, mask_bits, stream)
                   : launch_stream<1, false>(ctx, vp, view, loc4, vec, ld_vec, attn_io, out, ld_out, B, scale, nullptr, stream);
}
