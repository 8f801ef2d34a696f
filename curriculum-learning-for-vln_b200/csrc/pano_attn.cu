// Fused gather + feature dropout + soft-dot attention over the 36-view panorama
// (SoftDotAttention.forward units.py:107-118 on `_feature_variable` base.py:141-147 with
// EnvDropDecoder's env_drop policy.py:226-231), forward and backward in one kernel.
//
// Roofline class: HBM.  Algorithmic bytes: 36 x 2048 x 2 = 147 456 B per episode-step, read ONCE.
//
// v4.  One episode = one 2-CTA cluster; CTA `rank` owns feature columns [1024*rank, 1024*rank + 1024):
//   * a single SM ingests ~50 GB/s, so one CTA per episode (v2, v3) needed >= 3 us just to receive its
//     147 KB at the north-star batch of 64 episodes; two CTAs halve that and put 128 SMs to work;
//   * a column split needs no merge of 2048-wide accumulators: the two CTAs exchange only their 36
//     partial dot products (st.async into the peer's shared memory + its mbarrier transaction
//     count, 144 bytes), then run the softmax redundantly and each writes its own output columns;
//   * a half-panorama is 72 KB, so TWO units fit in shared memory: the next episode's rows (bulk-async
//     copies, one mbarrier per row) stream in while the current one is computed — the kernel is
//     persistent (cluster c takes episodes c, c + n_clusters, ...).
// Per unit:  phase 1  warp w owns rows w, w+12, w+24: keep-mask applied (row written back so phase 2
//                     never touches the mask), partial dot with the register-resident query half;
//            exchange + softmax (forward) or c_v = a_v (r_v - sum_u a_u r_u) (backward), one warp;
//            phase 2  column-parallel: thread group g (128 threads) owns rows 12g..12g+11, every thread
//                     one 16-byte column chunk: acc += w_v * row_v[chunk]; groups meet in shared memory.
// The 128 angle dimensions are 4 distinct values per view (misc.py:286-293): they ride along as 4
// scalars per row (loc4) against 4 group sums of the query.
//
//   forward :  logit_v = x~_v . q          a = softmax(logit)      out = sum_v a_v x~_v
//   backward:  r_v     = x~_v . d_out      c_v = a_v (r_v - a.r)   dq  = sum_v c_v x~_v
// with x~ = [ dropout(table row) | angle embedding ].
#include <cstdlib>

#include "common.cuh"

// optional phase stamps (VLN_PANO_STAMPS=k): CTA 0 / thread 0 records clock64 at phase boundaries of its k-th unit
__device__ unsigned long long g_pano_stamps[16];
#define PSTAMP(i)                                                                          \
  do {                                                                                     \
    if (dbg && blockIdx.x == 0 && tid == 0 && it == dbg - 1) g_pano_stamps[i] = (unsigned long long)clock64(); \
  } while (0)

namespace {

constexpr int kWarps = 12;
constexpr int kThreads = kWarps * 32;                    // 384
constexpr int kHalf = VLN_IMG / 2;                       // 1024 feature columns per CTA
constexpr int kRowBytes = kHalf * 2;                     // 2 048: half a table row
constexpr int kMaskBytes = kHalf / 8;                    // 128: packed keep-bits of half a row (vln_feature_mask_bits)
constexpr int kGroups = 3;                               // phase-2 thread groups (128 threads, 12 rows each)
constexpr int kRowsPerGroup = VLN_V / kGroups;           // 12
constexpr int kQ = kHalf + VLN_ANG;                      // floats of the query a CTA stages: its half + the angle part

__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// 4-byte store into the peer CTA's shared memory that counts on the peer's mbarrier
__device__ __forceinline__ void peer_st_async_f32(float* local_ptr, uint64_t* local_bar, uint32_t peer, float v) {
  uint32_t a = smem_u32(local_ptr), m = smem_u32(local_bar), ra, rm;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(peer));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rm) : "r"(m), "r"(peer));
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(ra),
               "r"(__float_as_uint(v)), "r"(rm)
               : "memory");
}

struct Smem {
  uint8_t rows[2][VLN_V * kRowBytes];       // two units: this CTA's half of the episode's panorama, bf16
  uint8_t mask[2][VLN_V * kMaskBytes];      // packed keep-bits (only with pre-generated masks)
  float qbuf[2][kQ];                        // query / d_out: [own 1024 columns | 128 angle columns]
  float locbuf[2][VLN_V * 4];               // loc4[cur_view] rows
  float attbuf[2][40];                      // saved attention (backward)
  float pmine[2][40];                       // partial dot products over my columns
  float ppeer[2][40];                       // ... the peer's, written through DSMEM
  float wv[40];                             // softmax weights (forward) / c_v (backward)
  float part[(kGroups - 1) * 128 * 8];      // phase-2 partial sums of groups 1, 2
  uint64_t full[2][VLN_V];
  uint64_t xbar[2];
};

__global__ void __launch_bounds__(kThreads, 1)
pano_attn_kernel(const __nv_bfloat16* __restrict__ table, const int32_t* __restrict__ vp,
                 const int32_t* __restrict__ view, const float* __restrict__ loc4, const float* __restrict__ vec,
                 float* __restrict__ attn_io, float* __restrict__ out, int B, int mode_in, float drop_p,
                 const uint64_t* __restrict__ rng, uint64_t call_off, int ld_vec, int ld_out,
                 const uint8_t* __restrict__ mask_bits, int gen_mask, int dbg, const __grid_constant__ ChainLink link) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rank = (int)cluster_ctarank();                 // which half of the feature columns
  const int cid = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int col0 = rank * kHalf;
  const int mode = mode_in & 1;
  // mode bit 2: latency-floor probe — the launch, the dependency wait, the dependent index load and ONE row's bulk copy
  // (one HBM round trip per CTA), nothing else: what a launch of this shape costs before a single useful byte moves
  const bool probe = (mode_in & 4) != 0;
  // mode bit 1: the indices, keep-bits and (backward) the saved attention were complete before the PRECEDING kernel
  // started (backward pass: they date from the forward pass) — the first unit's rows are then requested before the
  // programmatic-dependency wait, and the HBM round trip overlaps the predecessor's tail
  const bool early = (mode_in & 2) != 0;

  int it = 0;
  if (dbg && blockIdx.x == 0 && tid == 0) g_pano_stamps[0] = (unsigned long long)clock64();
  CHAIN_BEGIN(rng, 1 + 10 * mode);
  pdl_trigger();
  if (tid < 2 * VLN_V) mbar_init(&sm.full[0][0] + tid, 1);
  if (tid >= 96 && tid < 98) mbar_init(&sm.xbar[tid - 96], 1);
  fence_mbar_init();
  cluster_arrive();                                        // (waited for just before the first DSMEM store)
  // One unit per cluster (gen_mask): the keep-bits of this CTA's half panorama (the bytes vln_feature_mask_bits would
  // write: Philox block c of dense row ep*36 + v at byte (c % 32) * 4 + (c % 128) / 32 of the half row) are drawn HERE,
  // ahead of the dependency wait — the CTA is resident and idle while its predecessor runs, so the Philox rounds that
  // made the inline variant ALU-bound stay off the step's chain, and no bits cross HBM.
  auto draw_mask = [&]() {
    const uint64_t g_seed = rng[0], g_off = rng[1] + call_off;
    const uint32_t g_thr = drop_threshold(drop_p);
    for (int w = tid; w < VLN_V * kMaskBytes; w += kThreads) {
      const int v = w >> 7, wl = w & 127;
      const int c = rank * 128 + (wl & 3) * 32 + (wl >> 2);
      const Philox8 r = philox8(g_seed, g_off, ((uint64_t)cid * VLN_V + (uint64_t)v) * 256 + (uint64_t)c);
      uint32_t bits = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) bits |= (philox_keep(r, k, g_thr) ? 1u : 0u) << k;
      sm.mask[0][w] = (uint8_t)bits;
    }
  };
  if (gen_mask && cid < B && !early) draw_mask();           // (backward: after the rows have been requested, below)
  __syncthreads();
  if (!early) chain_wait_cta(link);                        // viewpoints, query and mask bits come from predecessors
  if (!early) CHAIN_MARK(2);
  if (dbg && blockIdx.x == 0 && tid == 0) g_pano_stamps[1] = (unsigned long long)clock64();

  // request row r of episode `ep` (this CTA's half, and its keep-bits) into unit buffer `ub`
  auto request_row = [&](int ep, int g, int r, int ub) {     // g = vp[ep], loaded one unit ahead
    uint64_t* bar = &sm.full[ub][r];
    mbar_expect_tx(bar, mask_bits ? kRowBytes + kMaskBytes : kRowBytes);
    bulk_g2s(sm.rows[ub] + (size_t)r * kRowBytes, table + ((size_t)g * VLN_V + r) * VLN_IMG + col0, kRowBytes, bar);
    if (mask_bits)
      bulk_g2s(sm.mask[ub] + (size_t)r * kMaskBytes,
               mask_bits + ((size_t)ep * VLN_V + r) * (2 * kMaskBytes) + (size_t)rank * kMaskBytes, kMaskBytes, bar);
  };
  const float scale = drop_p > 0.f ? 1.0f / (1.0f - drop_p) : 1.0f;
  const uint32_t thr = drop_threshold(drop_p);
  uint64_t seed = 0, offset = 0;
  if (drop_p > 0.f && !mask_bits && !gen_mask) {
    seed = rng[0];
    offset = rng[1] + call_off;
  }
  // stage the per-unit vectors (query half + angle part, angle table row, saved attention) of episode `e`
  auto prefetch_unit = [&](int e, int vw, int ub) {          // vw = view[e], loaded one unit ahead
    if (e < B) {
      const float* vr = vec + (size_t)e * ld_vec;
      for (int c = tid; c < kQ / 4; c += kThreads) {
        const int src = c < kHalf / 4 ? col0 / 4 + c : VLN_IMG / 4 + (c - kHalf / 4);
        cp_async16(&sm.qbuf[ub][c * 4], vr + src * 4);
      }
      if (tid < VLN_V) cp_async16(&sm.locbuf[ub][tid * 4], loc4 + (size_t)vw * (VLN_V * 4) + tid * 4);
      if (mode == 1 && tid >= 64 && tid < 64 + VLN_V / 4)
        cp_async16(&sm.attbuf[ub][(tid - 64) * 4], attn_io + (size_t)e * VLN_V + (tid - 64) * 4);
    }
    cp_async_commit();
  };

  // viewpoint / view indices are fetched one unit ahead of their use, so no warp ever stalls on that
  // dependent load inside the unit loop (it cost ~2 us per unit at large B)
  int g_next = 0, vw_next = 0;                             // indices of the NEXT unit (ep + n_clusters)
  {
    const int g0 = cid < B ? __ldg(vp + cid) : 0, vw0 = cid < B ? __ldg(view + cid) : 0;
    if (cid + n_clusters < B) {
      g_next = __ldg(vp + cid + n_clusters);
      vw_next = __ldg(view + cid + n_clusters);
    }
    // every warp requests the three rows it will consume in phase 1 (one warp issuing all 36-72 bulk copies
    // serialised ~30 cycles apiece: the first vectors were ready 1 us later with keep-bits than without)
    if (probe) {
      if (cid < B && tid == 0) {
        request_row(cid, g0, 0, 0);
        mbar_wait(&sm.full[0][0], 0);
        attn_io[(size_t)cid * VLN_V] = (float)sm.rows[0][0];
      }
      cluster_wait();
      chain_signal_cta(link);
      return;
    }
    if (dbg && blockIdx.x == 0 && tid == 0 && g0 >= 0) g_pano_stamps[9] = (unsigned long long)clock64();    // index in hand
    if (cid < B && lane < VLN_V / kWarps) request_row(cid, g0, warp + lane * kWarps, 0);
    if (dbg && blockIdx.x == 0 && tid == 0) g_pano_stamps[10] = (unsigned long long)clock64();             // rows requested
    if (gen_mask && cid < B && early) draw_mask();         // while the rows are in flight
    if (early) chain_wait_cta(link);                       // the query / gradient vector comes from the predecessor
    if (early) CHAIN_MARK(2);
    prefetch_unit(cid, vw0, 0);
  }

  for (int ep = cid; ep < B; ep += n_clusters, ++it) {
    const int ub = it & 1;
    const uint32_t ph = (uint32_t)(it >> 1) & 1u;          // every barrier of a unit buffer completes once per two units
    PSTAMP(8);
    const int next_ep = ep + n_clusters;
    // the other unit buffer was last read in the previous iteration (which ended with a __syncthreads)
    const int g_req = g_next, vw_req = vw_next;
    if (next_ep + n_clusters < B) {                        // consumed one iteration from now
      g_next = __ldg(vp + next_ep + n_clusters);
      vw_next = __ldg(view + next_ep + n_clusters);
    }
    if (next_ep < B && lane < VLN_V / kWarps) {
      fence_proxy_async();                                 // order those generic-proxy accesses before the async writes
      request_row(next_ep, g_req, warp + lane * kWarps, ub ^ 1);
    }
    if (tid == 0) mbar_expect_tx(&sm.xbar[ub], VLN_V * 4); // the peer's 36 partial dot products of this unit
    cp_async_wait_all();
    __syncthreads();                                       // this unit's vectors are in place for every warp
    prefetch_unit(next_ep, vw_req, ub ^ 1);                // overlaps with this unit's work
    PSTAMP(2);
    if (it == 0) cluster_wait();                           // the peer's barriers exist before anything is sent to it

    // ------------------------------ phase 1: partial logits over my columns ------------------------------
    {
      float q[32];                                         // q[j*8+e] = vec[col0 + j*256 + lane*8 + e]
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 a = *reinterpret_cast<const float4*>(&sm.qbuf[ub][j * 256 + lane * 8]);
        const float4 b = *reinterpret_cast<const float4*>(&sm.qbuf[ub][j * 256 + lane * 8 + 4]);
        q[j * 8 + 0] = a.x; q[j * 8 + 1] = a.y; q[j * 8 + 2] = a.z; q[j * 8 + 3] = a.w;
        q[j * 8 + 4] = b.x; q[j * 8 + 5] = b.y; q[j * 8 + 6] = b.z; q[j * 8 + 7] = b.w;
      }
      // angle group sums of the query: qa[k] = sum_i vec[2048 + 32k + i]; rank 0 adds the angle term of the logit
      float qa0 = 0.f, qa1 = 0.f, qa2 = 0.f, qa3 = 0.f;
      if (rank == 0) {
        const float4 a = *reinterpret_cast<const float4*>(&sm.qbuf[ub][kHalf + lane * 4]);
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
        mbar_wait(&sm.full[ub][v], ph);
        uint4* rowp = reinterpret_cast<uint4*>(sm.rows[ub] + (size_t)v * kRowBytes);
        uint4 x[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = rowp[j * 32 + lane];
        if (drop_p > 0.f) {
          if (mask_bits || gen_mask) {
            // pre-generated keep-bits: byte j of this lane's 4-byte group covers the 8 features of x[j]
            const uint32_t mb = *reinterpret_cast<const uint32_t*>(sm.mask[ub] + (size_t)v * kMaskBytes + lane * 4);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t bits = (mb >> (j * 8)) & 0xFFu;
              uint32_t* w = reinterpret_cast<uint32_t*>(&x[j]);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                w[k] &= (((bits >> (2 * k)) & 1u) * 0x0000FFFFu) | (((bits >> (2 * k + 1)) & 1u) * 0xFFFF0000u);
            }
          } else {
            const uint64_t e0 = ((((uint64_t)ep * VLN_V + (uint64_t)v) * VLN_IMG) >> 3) + (uint64_t)(rank * (kHalf / 8));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const Philox8 r = philox8(seed, offset, e0 + (uint64_t)(j * 32 + lane));
              uint32_t* w = reinterpret_cast<uint32_t*>(&x[j]);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                w[k] &= (philox_keep(r, 2 * k, thr) ? 0x0000FFFFu : 0u) | (philox_keep(r, 2 * k + 1, thr) ? 0xFFFF0000u : 0u);
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) rowp[j * 32 + lane] = x[j];     // phase 2 reads the masked row
        }
        float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;         // four chains: no 32-deep FMA dependency
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          d0 = fmaf(bf16lo(x[j].x), q[j * 8 + 0], d0); d1 = fmaf(bf16hi(x[j].x), q[j * 8 + 1], d1);
          d2 = fmaf(bf16lo(x[j].y), q[j * 8 + 2], d2); d3 = fmaf(bf16hi(x[j].y), q[j * 8 + 3], d3);
          d0 = fmaf(bf16lo(x[j].z), q[j * 8 + 4], d0); d1 = fmaf(bf16hi(x[j].z), q[j * 8 + 5], d1);
          d2 = fmaf(bf16lo(x[j].w), q[j * 8 + 6], d2); d3 = fmaf(bf16hi(x[j].w), q[j * 8 + 7], d3);
        }
        float dot = warp_sum((d0 + d1) + (d2 + d3)) * scale;
        if (rank == 0) {
          const float4 lv = *reinterpret_cast<const float4*>(&sm.locbuf[ub][v * 4]);   // this view's angle values
          dot += lv.x * qa0 + lv.y * qa1 + lv.z * qa2 + lv.w * qa3;
        }
        if (lane == 0) {
          sm.pmine[ub][v] = dot;
          peer_st_async_f32(&sm.ppeer[ub][v], &sm.xbar[ub], (uint32_t)(rank ^ 1), dot);
        }
      }
    }
    PSTAMP(3);
    __syncthreads();
    PSTAMP(4);

    // ------------------------------ exchange, softmax / backward coefficients ------------------------------
    if (warp == 0) {
      mbar_wait(&sm.xbar[ub], ph);                         // the peer's partials have landed
      const bool two = lane < VLN_V - 32;
      const float* lb = sm.locbuf[ub];
      const int v1 = two ? 32 + lane : lane;
      // both CTAs add the two partials in the same order, so they agree bit for bit
      const float* p0 = rank == 0 ? sm.pmine[ub] : sm.ppeer[ub];
      const float* p1 = rank == 0 ? sm.ppeer[ub] : sm.pmine[ub];
      const float x0 = p0[lane] + p1[lane];
      const float x1 = two ? p0[v1] + p1[v1] : -INFINITY;
      float w0, w1;
      if (mode == 0) {
        const float m = warp_max(fmaxf(x0, x1));
        const float e0 = __expf(x0 - m), e1 = two ? __expf(x1 - m) : 0.f;
        const float inv = 1.0f / warp_sum(e0 + e1);
        w0 = e0 * inv;
        w1 = e1 * inv;
        if (rank == 0) {
          attn_io[(size_t)ep * VLN_V + lane] = w0;
          if (two) attn_io[(size_t)ep * VLN_V + 32 + lane] = w1;
        }
      } else {
        const float a0 = sm.attbuf[ub][lane], a1 = two ? sm.attbuf[ub][32 + lane] : 0.f;
        const float rbar = warp_sum(a0 * x0 + (two ? a1 * x1 : 0.f));
        w0 = a0 * (x0 - rbar);
        w1 = two ? a1 * (x1 - rbar) : 0.f;
      }
      sm.wv[lane] = w0;
      if (two) sm.wv[32 + lane] = w1;
      if (rank == 0) {                                     // the 128 angle dimensions of the output: 4 values, each x32
        __syncwarp();
        const int k = lane >> 3, part8 = lane & 7;         // 8 lanes per angle value, 4-5 views each
        float g = 0.f;
        for (int v = part8; v < VLN_V; v += 8) g = fmaf(sm.wv[v], lb[v * 4 + k], g);
        g += __shfl_xor_sync(0xffffffffu, g, 1);
        g += __shfl_xor_sync(0xffffffffu, g, 2);
        g += __shfl_xor_sync(0xffffffffu, g, 4);
        float* orow = out + (size_t)ep * ld_out + VLN_IMG + k * 32;
#pragma unroll
        for (int i = 0; i < 4; ++i) orow[part8 * 4 + i] = g;
      }
    }
    __syncthreads();
    PSTAMP(5);

    // ------------------------------ phase 2: weighted sum, column-parallel ------------------------------
    {
      const int g = tid >> 7, t = tid & 127;               // group g: rows 12g..12g+11; 16-byte column chunk t
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll 4
      for (int i = 0; i < kRowsPerGroup; ++i) {
        const int v = g * kRowsPerGroup + i;
        const float w = sm.wv[v];
        const uint4 xa = reinterpret_cast<const uint4*>(sm.rows[ub] + (size_t)v * kRowBytes)[t];
        acc[0] = fmaf(w, bf16lo(xa.x), acc[0]); acc[1] = fmaf(w, bf16hi(xa.x), acc[1]);
        acc[2] = fmaf(w, bf16lo(xa.y), acc[2]); acc[3] = fmaf(w, bf16hi(xa.y), acc[3]);
        acc[4] = fmaf(w, bf16lo(xa.z), acc[4]); acc[5] = fmaf(w, bf16hi(xa.z), acc[5]);
        acc[6] = fmaf(w, bf16lo(xa.w), acc[6]); acc[7] = fmaf(w, bf16hi(xa.w), acc[7]);
      }
      PSTAMP(6);
      if (g > 0) {
        float* p = sm.part + ((size_t)(g - 1) * 128 + t) * 8;
        *reinterpret_cast<float4*>(p) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
      }
      __syncthreads();
      if (g == 0) {
#pragma unroll
        for (int gg = 0; gg < kGroups - 1; ++gg) {
          const float* p = sm.part + ((size_t)gg * 128 + t) * 8;
          const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
          acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
          acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
        }
        float* orow = out + (size_t)ep * ld_out + col0 + t * 8;
        reinterpret_cast<float4*>(orow)[0] = make_float4(acc[0] * scale, acc[1] * scale, acc[2] * scale, acc[3] * scale);
        reinterpret_cast<float4*>(orow)[1] = make_float4(acc[4] * scale, acc[5] * scale, acc[6] * scale, acc[7] * scale);
      }
    }
    __syncthreads();                                       // unit buffer `ub`, wv, part are free for reuse
    PSTAMP(7);
  }
  // No CTA exits while its peer could still store into it: every unit's exchange was waited for above, and a
  // cluster without any unit (cid >= B) exchanges nothing; it only completes the initial cluster barrier.
  if (it == 0) cluster_wait();
  chain_signal_cta(link);
  CHAIN_MARK(3);
}

}  // namespace

extern "C" int vln_pano_attn_ld(const vln_ctx* ctx, const int32_t* vp, const int32_t* view, const float* loc4,
                                const float* vec, int ld_vec, float* attn_io, const float* fwd_out, int ld_fwd,
                                float* out, int ld_out, int B, int mode, float drop_p, const uint64_t* rng,
                                uint64_t call_off, const uint8_t* mask_bits, int split, void* stream) {
  VLN_REQUIRE(ctx && vp && view && loc4 && vec && attn_io && out && B > 0, "bad arguments");
  VLN_REQUIRE(split == 1 || split == 2 || split == 4, "split (kernel variant) must be 1 = automatic, 2 = cluster, 4 = streaming");
  VLN_REQUIRE(mode >= 0 && mode <= 7, "mode must be 0 (forward) or 1 (backward), + 2 = indices complete before the predecessor, + 4 = latency-floor probe");
  VLN_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "drop_p out of range");
  VLN_REQUIRE(drop_p == 0.f || rng || mask_bits, "dropout needs an rng state or pre-generated keep-bits");
  VLN_REQUIRE(!mask_bits || (drop_p > 0.f && ((uintptr_t)mask_bits & 15) == 0), "mask_bits: 16-byte aligned, with drop_p > 0");
  VLN_REQUIRE(ld_vec >= VLN_F && ld_out >= VLN_F, "row strides must be >= 2176");
  VLN_REQUIRE(ld_vec % 4 == 0 && ld_out % 4 == 0 && ((uintptr_t)vec & 15) == 0 && ((uintptr_t)out & 15) == 0,
              "rows must be 16-byte aligned");
  (void)fwd_out;
  (void)ld_fwd;
  // Two kernels serve this entry point: the 2-CTA-cluster kernel below minimises the latency of a launch with few
  // episodes (the rollout's B = 64), the streaming kernel (pano_stream.cu) maximises bytes in flight for many.
  static const int stream_min_b = getenv("VLN_PANO_STREAM_MIN_B") ? atoi(getenv("VLN_PANO_STREAM_MIN_B")) : 128;
  const bool stream_ok = drop_p == 0.f || mask_bits;        // the streaming kernel has no inline Philox
  if (stream_ok && !(mode & 4) && (split == 4 || (split == 1 && B >= stream_min_b))) {
    VLN_CHECK_CUDA(vln_pano_stream_launch(ctx, vp, view, loc4, vec, ld_vec, attn_io, out, ld_out, B, mode & 1, drop_p, mask_bits,
                                          (cudaStream_t)stream));
    return 0;
  }
  static bool configured = false;
  if (!configured) {
    VLN_CHECK_CUDA(cudaFuncSetAttribute(pano_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
    configured = true;
  }
  const int max_clusters = ctx->num_sms / 2;
  const int clusters = B < max_clusters ? B : max_clusters;
  // dropout without pre-generated bits: one unit per cluster draws them in shared memory ahead of the dependency wait
  const int gen_mask = (drop_p > 0.f && !mask_bits && B <= max_clusters) ? 1 : 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = sizeof(Smem);
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;          // the two column halves of an episode
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = vln_pdl_enabled() ? 2 : 1;
  const ChainLink link = vln_chain_link((cudaStream_t)stream, (unsigned int)(2 * clusters));
  VLN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, pano_attn_kernel, ctx->table, vp, view, loc4, vec, attn_io, out, B, mode, drop_p, rng,
                                    call_off, ld_vec, ld_out, mask_bits, gen_mask, (getenv("VLN_PANO_STAMPS") ? atoi(getenv("VLN_PANO_STAMPS")) : 0),
                                    link));
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
