// Skinny linear layers on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
//   y[m, n] += sum_k x[m, k] * w[n, k]  (+ bias[n])        m < M <= 128 (a batch of episodes),
//                                                           n < N (gate rows / projection outputs)
// used for the decoder LSTMCell gates (policy.py:53,159,238), the 512<->2176 attention / candidate
// projections (units.py:107, policy.py:199-206), linear_in/linear_out, and — with the transposed
// weight copies — for their input gradients.
//
// Swap-AB: the WEIGHT rows sit on MMA-M (tiles of 128), the batch on MMA-N (64 or 128), so a 64-row
// batch still fills the 128-lane datapath; split-K spreads a layer over ~all 148 SMs.  Every CTA parks
// its partial accumulator tile in shared memory (transposed, so a thread owns 4 consecutive outputs of
// one batch row) and adds it to y with 16-byte vector reductions (REDG.ADD.F32x4); y is zero-filled
// first unless the caller accumulates into it.  An alternative merge — the splits of a tile as a
// thread-block cluster reducing through distributed shared memory, plain stores — is kept behind
// VLN_GEMM_VARIANT=c8|c4|c2 for the record: on B200 it measured 1-7 us slower per launch (cluster
// barrier + DSMEM reads cost ~3 us, and 17 clusters of 8 CTAs need two waves).
//
// Precision: bf16x3.  Weights are pre-split once per optimiser step into bf16 hi + lo
// (vln_split_bf16), activations are split on the fly while they are staged; three MMAs
// (hi.hi + hi.lo + lo.hi, fp32 accumulate in TMEM) reproduce the fp32 product to ~2^-16 relative,
// which keeps the rollout inside the 1e-3 logit/loss bar that a plain bf16 or tf32 GEMM would miss.
//
// Warp roles (256 threads): warp 0 = TMA producer (weight tiles as 128B-swizzled bf16 boxes + the raw fp32
// activation tile), warp 1 = MMA issuer (one elected lane), warps 4-7 = activation converters during the
// main loop (shared fp32 -> bf16 hi/lo, written with the 128B swizzle by hand) and epilogue afterwards
// (tcgen05.ld -> shared memory); all 8 warps then take part in the cluster reduction.  4-stage mbarrier
// ring, accumulator in TMEM.
#include <cstdlib>
#include <cstring>
#include <cudaTypedefs.h>

#include "common.cuh"

int vln_make_tmap_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                     uint32_t box_cols, uint32_t box_rows, int swizzle128);
int vln_make_tmap_2d_f32(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                         uint32_t box_cols, uint32_t box_rows);

// ---- optional phase stamps (VLN_GEMM_STAMPS=1): CTA (0,0) records clock64 at 9 phase boundaries + globaltimer ----
constexpr int kStampSlots = 1024;
__device__ unsigned long long g_stamps[kStampSlots * 12];
__device__ unsigned int g_stamp_n;
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define STAMP(i)                                                                    \
  do {                                                                              \
    if (dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_stamps[(slot % kStampSlots) * 12 + (i)] = (unsigned long long)clock64(); \
  } while (0)

namespace {

constexpr int kBK = 64;                 // k-block: 64 bf16 = one 128-byte swizzle row
constexpr int kTileN = 128;             // weight rows per CTA (MMA M)
constexpr int kThreads = 256;

// ST ring stages: 3 (MP = 64) / 2 (MP = 128) for the deep-K layers; launches whose CTAs run only one or two k-blocks
// (most of the step's GEMMs after split-K) take 1 or 2 stages = 65 / 130 KB instead of 193 KB of shared memory, so
// that the CTAs of the NEXT kernel of the chain become resident while this one still runs and programmatic
// dependent launch really overlaps their prologue + weight-tile requests with it.
template <int MP, int ST>
struct Cfg {
  static constexpr int kStages = ST;
  static constexpr int kABytes = kTileN * kBK * 2;   // 16 KB per hi / lo weight tile
  static constexpr int kBBytes = MP * kBK * 2;       // 8 / 16 KB per hi / lo activation tile
  static constexpr int kXBytes = MP * kBK * 4;       // 16 / 32 KB raw fp32 activation tile (TMA destination)
  static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes + kXBytes;
  static constexpr int kSmem = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

__device__ __forceinline__ uint64_t make_sdesc(uint32_t smem_addr) {
  // K-major, SWIZZLE_128B: 8-row groups of 128-byte rows, SBO = 1024 B, LBO = 1 (unused), version 1 (sm_100)
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// (called by a whole converged warp with warp-uniform operands; one elected lane issues: with the operands in
// uniform registers the issue loop is UIADD3 + UTCHMMA, instead of ELECT + four R2UR.BROADCAST per MMA)
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// one row of a CTA's partial tile, shared -> global, added (split-K merge) or stored by the bulk-copy engine: the L2
// performs the 512-byte row as whole-sector operations instead of 32 separate 16-byte REDs from the LSU
__device__ __forceinline__ void bulk_reduce_add_f32(float* gdst, const float* ssrc, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_store_f32(float* gdst, const float* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_wait_all() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {   // round-to-nearest-even, a in the low half
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}

// ---- optional tile epilogue (the per-element glue of the decoder step, folded into its producer GEMM) ----
// The split-K partials of a tile meet in y through reductions, so no CTA owns a finished output.  With an epilogue
// the `splits` CTAs of a weight tile — all resident: the grid is at most one CTA per SM — count themselves in on a
// per-tile counter after their reductions (fence + atomic), spin until the tile is complete, and then each applies
// the epilogue to 1/splits of the tile's [M x 128] outputs, read back from L2.  That replaces a separate grid-wide
// launch on the step's latency chain (state_fwd / state_bwd, csrc/step.cu: same arithmetic, same Philox streams).
struct Epi {
  int kind;                      // 0 none; 1 state_fwd on y = linear_out pre-activation; 2 state_bwd on y = d_hq_next
  unsigned int* ctr;             // [2 * tiles] arrive / depart counters, zero between launches
  float p;
  const uint64_t* rng;
  unsigned long long off_q, off_c;
  float* xh_next; int ld_xh; float* hq_next; float* hc_cur;                                  // kind 1 outputs
  const float* d_hc; const float* d_xh_next; int ld_dxh; const float* htilde; int ld_h;      // kind 2 inputs
  float* d_src;                                                                               // kind 2 output
};

__device__ __forceinline__ void ldcg8(const float* p, float (&v)[8]) {
  const float4 a = __ldcg(reinterpret_cast<const float4*>(p)), b = __ldcg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

__device__ __forceinline__ void tile_epilogue(const Epi& e, const float* y, int ldy, int M, int N, int tile, int split,
                                              int splits, int tid) {
  unsigned int* arrive = e.ctr + 2 * tile;
  unsigned int* depart = arrive + 1;
  __threadfence();                                           // this thread's reductions into y are ordered before the count
  __syncthreads();
  if (tid == 0) {
    atomicAdd(arrive, 1u);
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(arrive) : "memory");
    } while (seen < (unsigned int)splits);
  }
  __syncthreads();
  __threadfence();
  const int n_items = M * (kTileN / 8);                      // items of 8 consecutive outputs (one Philox block)
  const int chunk = (n_items + splits - 1) / splits;
  const int i0 = split * chunk, i1 = min(n_items, i0 + chunk);
  const bool on = e.p > 0.f;
  const uint32_t thr = drop_threshold(e.p);
  const float sc = on ? 1.0f / (1.0f - e.p) : 1.0f;
  const uint64_t seed = on ? e.rng[0] : 0, base = on ? e.rng[1] : 0;
  for (int it = i0 + tid; it < i1; it += kThreads) {
    const int m = it / (kTileN / 8), g = it - m * (kTileN / 8);
    const int n = tile * kTileN + g * 8;
    if (n >= N) continue;
    float v[8], kq[8], kc[8];
    ldcg8(y + (size_t)m * ldy + n, v);
    const uint64_t blk = ((uint64_t)m * (uint64_t)N + (uint64_t)n) >> 3;
    if (on) {
      const Philox8 rq = philox8(seed, base + e.off_q, blk), rc = philox8(seed, base + e.off_c, blk);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        kq[j] = philox_keep(rq, j, thr) ? sc : 0.f;
        kc[j] = philox_keep(rc, j, thr) ? sc : 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) kq[j] = kc[j] = 1.f;
    }
    if (e.kind == 1) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = tanhf(v[j]);
      if (e.xh_next) st8(e.xh_next + (size_t)m * e.ld_xh + n, v);
      if (e.hq_next) {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = v[j] * kq[j];
        st8(e.hq_next + (size_t)m * N + n, o);
      }
      if (e.hc_cur) {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = v[j] * kc[j];
        st8(e.hc_cur + (size_t)m * N + n, o);
      }
    } else {
      float a[8], b[8], h[8], o[8];
      ldcg8(e.d_hc + (size_t)m * N + n, a);
      ldcg8(e.d_xh_next + (size_t)m * e.ld_dxh + n, b);
      ldcg8(e.htilde + (size_t)m * e.ld_h + n, h);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (a[j] * kc[j] + b[j] + v[j] * kq[j]) * (1.f - h[j] * h[j]);
      st8(e.d_src + (size_t)m * N + n, o);
    }
  }
  __syncthreads();
  if (tid == 0) {
    const unsigned int d = atomicAdd(depart, 1u);
    if (d == (unsigned int)splits - 1u) {                    // everyone has left the spin: counters back to zero
      *arrive = 0u;
      *depart = 0u;
    }
  }
}

template <int MP, int ST>
__global__ void __launch_bounds__(kThreads, 1)
linear_bf16x3_kernel(const __grid_constant__ CUtensorMap tm_hi0, const __grid_constant__ CUtensorMap tm_lo0,
                     const __grid_constant__ CUtensorMap tm_x0, const __grid_constant__ CUtensorMap tm_hi1,
                     const __grid_constant__ CUtensorMap tm_lo1, const __grid_constant__ CUtensorMap tm_x1, int M, int N,
                     int K, const float* __restrict__ bias, float* __restrict__ y0, float* __restrict__ y1, int ldy,
                     int splits, int accumulate, int mode, int dbg, int m_tiles, const __grid_constant__ Epi epi,
                     const __grid_constant__ ChainLink link) {
  // blockIdx.z selects one of two independent problems of identical shape (e.g. the candidate projection of
  // step t and the visual-attention query of step t+1, which both only wait for h~_t)
  const CUtensorMap& tm_hi = blockIdx.z == 0 ? tm_hi0 : tm_hi1;
  const CUtensorMap& tm_lo = blockIdx.z == 0 ? tm_lo0 : tm_lo1;
  const CUtensorMap& tm_x = blockIdx.z == 0 ? tm_x0 : tm_x1;
  float* __restrict__ y = blockIdx.z == 0 ? y0 : y1;
  using C = Cfg<MP, ST>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // swizzle atoms: 1 KB aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + C::kStages * C::kStageBytes);
  uint64_t* full_w = bars;                       // TMA bytes landed (weight tiles + raw activation tile)
  uint64_t* full_x = bars + C::kStages;          // activation tile converted (4 warp arrivals)
  uint64_t* empty = bars + 2 * C::kStages;       // MMAs that read the stage have completed
  uint64_t* acc_done = bars + 3 * C::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * C::kStages + 1);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);           // warp-uniform by construction
  __shared__ unsigned int slot_s;
  unsigned int slot = 0;
  if (dbg && blockIdx.x == 0 && blockIdx.y == 0) {
    if (tid == 0) {
      slot_s = atomicAdd(&g_stamp_n, 1u);
      g_stamps[(slot_s % kStampSlots) * 12 + 10] = gtimer();
      g_stamps[(slot_s % kStampSlots) * 12 + 0] = (unsigned long long)clock64();
    }
    __syncthreads();
    slot = slot_s;
  }
  // skinny launches: blockIdx.y = split-K slice of one weight tile (the `splits` CTAs of a tile may form a cluster);
  // tall launches (m_tiles > 1, splits == 1): blockIdx.y = block of MP activation rows, plain stores into y
  const int tile = blockIdx.x, split = m_tiles > 1 ? 0 : (int)blockIdx.y;
  const int m0 = m_tiles > 1 ? (int)blockIdx.y * MP : 0;
  const int nkb = K / kBK;
  const int kb0 = (int)((long long)split * nkb / splits), kb1 = (int)((long long)(split + 1) * nkb / splits);
  const int n_iter = kb1 - kb0;

  pdl_trigger();
  if (tid == 0) {
    tma_prefetch_desc(&tm_hi);
    tma_prefetch_desc(&tm_lo);
    tma_prefetch_desc(&tm_x);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full_w[s], 1);
      mbar_init(&full_x[s], 4);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_done, 1);
    fence_mbar_init();
  }
  if (warp == 0) {   // TMEM accumulator: MP fp32 columns x 128 lanes
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(MP)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  if (tid == 0) STAMP(1);

  if (n_iter > 0) {
    if (warp == 0) {
      // ---------------- TMA producer: weight tiles (hi, lo) ----------------
      if (lane == 0) {
        for (int it = 0; it < n_iter; ++it) {
          const int s = it % C::kStages;
          const uint32_t ph = (it / C::kStages) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          uint8_t* st = base + (size_t)s * C::kStageBytes;
          mbar_expect_tx(&full_w[s], 2 * C::kABytes + C::kXBytes);
          tma_load_2d(st, &tm_hi, &full_w[s], (kb0 + it) * kBK, tile * kTileN);
          tma_load_2d(st + C::kABytes, &tm_lo, &full_w[s], (kb0 + it) * kBK, tile * kTileN);
          // the weight tiles were written before this chain of kernels started; the activations come from the
          // predecessor: everything downstream (conversion, MMA, epilogue) is ordered after this wait
          if (it == 0) {
            chain_wait_thread(link);
            if (dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_stamps[(slot % kStampSlots) * 12 + 9] = gtimer();
          }
          tma_load_2d(st + 2 * C::kABytes + 2 * C::kBBytes, &tm_x, &full_w[s], (kb0 + it) * kBK, m0);
        }
      }
    } else if (warp == 1) {
      // ---------------- MMA issuer (whole warp converged, one elected lane issues) ----------------
      // instruction descriptor: D = F32, A = B = BF16, both K-major, N = MP, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(MP >> 3) << 17) | ((uint32_t)(kTileN >> 4) << 24);
      const uint32_t base_addr = smem_u32(base);
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % C::kStages;
        const uint32_t ph = (it / C::kStages) & 1u;
        mbar_wait(&full_w[s], ph);
        if (it == 0 && lane == 0) STAMP(2);
        mbar_wait(&full_x[s], ph);
        if (it == 0 && lane == 0) STAMP(3);
        tc_fence_after();
        const uint32_t a_hi = base_addr + (uint32_t)s * C::kStageBytes, a_lo = a_hi + C::kABytes;
        const uint32_t b_hi = a_lo + C::kABytes, b_lo = b_hi + C::kBBytes;
        const uint64_t da_hi = make_sdesc(a_hi), da_lo = make_sdesc(a_lo), db_hi = make_sdesc(b_hi), db_lo = make_sdesc(b_lo);
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) {
          const uint64_t ko = (uint64_t)(k * 2);              // 16 bf16 = 32 bytes along the swizzled row, in 16-byte units
          umma_bf16(tmem_d, da_hi + ko, db_hi + ko, idesc, (it | k) != 0);
          umma_bf16(tmem_d, da_hi + ko, db_lo + ko, idesc, 1u);
          umma_bf16(tmem_d, da_lo + ko, db_hi + ko, idesc, 1u);
        }
        umma_commit(&empty[s]);                              // frees the stage when these MMAs retire
      }
      umma_commit(acc_done);
      __syncwarp();
    } else if (warp >= 4) {
      // ---------------- activation converters: raw fp32 tile (TMA) -> bf16 hi/lo, 128B-swizzled rows ----------------
      // A row of the raw tile is 64 floats = 16 float4; lane l of a warp takes float4 (l % 16) of row (l / 16),
      // so shared-memory reads are conflict-free and nothing on this path waits for L2.
      const int t = tid - 128;                               // 0..127
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % C::kStages;
        const uint32_t ph = (it / C::kStages) & 1u;
        mbar_wait(&full_w[s], ph);
        uint8_t* bh = base + (size_t)s * C::kStageBytes + 2 * C::kABytes;
        uint8_t* bl = bh + C::kBBytes;
        const float4* raw = reinterpret_cast<const float4*>(bl + C::kBBytes);
#pragma unroll
        for (int i = 0; i < MP / 8; ++i) {
          const int idx = i * 128 + t;                       // float4 index in the [MP][16] tile
          const int r = idx >> 4, q = idx & 15;
          const float4 f = raw[idx];
          const float ax = __bfloat162float(__float2bfloat16_rn(f.x)), ay = __bfloat162float(__float2bfloat16_rn(f.y));
          const float az = __bfloat162float(__float2bfloat16_rn(f.z)), aw = __bfloat162float(__float2bfloat16_rn(f.w));
          const uint32_t off = (uint32_t)r * 128u + (uint32_t)((((q >> 1) ^ (r & 7)) << 4) + ((q & 1) << 3));
          *reinterpret_cast<uint2*>(bh + off) = make_uint2(pack_bf16(f.x, f.y), pack_bf16(f.z, f.w));
          *reinterpret_cast<uint2*>(bl + off) = make_uint2(pack_bf16(f.x - ax, f.y - ay), pack_bf16(f.z - az, f.w - aw));
        }
        fence_proxy_async();                                 // generic-proxy stores -> visible to the MMA (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_x[s]);
      }
      // ---------------- epilogue, part 1: TMEM -> this CTA's partial tile red[m][n] in shared memory ----------------
      // (the ring is free: acc_done fires after every MMA that read it has retired)
      mbar_wait(acc_done, 0);
      if (tid == 128) STAMP(4);
      tc_fence_after();
      const int wq = warp & 3;                               // TMEM lane quarter this warp may access
      float* red = reinterpret_cast<float*>(base);
#pragma unroll
      for (int cb = 0; cb < MP / 32; ++cb) {
        uint32_t v[32];
        tmem_ld32(tmem_d + ((uint32_t)(wq * 32) << 16) + (uint32_t)(cb * 32), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) red[(cb * 32 + j) * kTileN + wq * 32 + lane] = __uint_as_float(v[j]);
      }
      fence_proxy_async();                                   // the bulk-copy engine (async proxy) reads this tile below
    }
  }
  // ---------------- epilogue, part 2: cluster reduction through distributed shared memory ----------------
  if (tid == 128) STAMP(5);
  tc_fence_before();
  __syncthreads();
  if (tid == 0) STAMP(6);
  if (mode == 3) {
    // (variant) split-K merge by the bulk-copy engine: one cp.reduce.async.bulk (add.f32) per row of the partial tile
    // instead of the per-thread vector reductions below (which take 2-3 us of every launch in the step chain)
    float* red = reinterpret_cast<float*>(base);
    if (bias != nullptr && split == 0 && n_iter > 0) {       // (CTA-uniform) the bias joins the first split's tile
      for (int item = tid; item < MP * (kTileN / 4); item += kThreads) {
        const int m = item / (kTileN / 4), c = (item - m * (kTileN / 4)) * 4;
        const int n = tile * kTileN + c;
        if (n >= N) continue;
        const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + n));
        float4* p = reinterpret_cast<float4*>(red + m * kTileN + c);
        float4 a = *p;
        a.x += bv.x; a.y += bv.y; a.z += bv.z; a.w += bv.w;
        *p = a;
      }
      fence_proxy_async();
      __syncthreads();
    }
    const int rows_here = min(MP, M - m0);
    const int n0 = tile * kTileN;
    const uint32_t row_bytes = (uint32_t)min(kTileN, N - n0) * 4u;
    if (n_iter > 0 && tid < rows_here) {
      float* dst = y + (size_t)(m0 + tid) * ldy + n0;
      if (m_tiles > 1 && !accumulate) bulk_store_f32(dst, red + tid * kTileN, row_bytes);   // the only writer of this block
      else bulk_reduce_add_f32(dst, red + tid * kTileN, row_bytes);
      if (epi.kind != 0) {
        bulk_commit_wait_all();                              // complete (not merely read): the tile epilogue re-reads y
      } else {                                               // the source rows have been read; grid completion covers the rest
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
    }
    if (epi.kind != 0) tile_epilogue(epi, y, ldy, M, N, tile, split, splits, tid);
  } else if (mode == 2) {
    // partial tiles meet in y through 16-byte vector reductions (REDG.ADD.F32x4); VLN_GEMM_BULK=0 selects this path
    const float* red = reinterpret_cast<const float*>(base);
    for (int item = tid; item < MP * (kTileN / 4); item += kThreads) {
      const int m = item / (kTileN / 4), c = (item - m * (kTileN / 4)) * 4;
      const int n = tile * kTileN + c;
      if (m0 + m >= M || n >= N) continue;
      float4 acc = *reinterpret_cast<const float4*>(red + m * kTileN + c);
      if (bias != nullptr && split == 0) {
        const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + n));
        acc.x += bv.x; acc.y += bv.y; acc.z += bv.z; acc.w += bv.w;
      }
      float* dst = y + (size_t)(m0 + m) * ldy + n;
      if (m_tiles > 1 && !accumulate) {
        *reinterpret_cast<float4*>(dst) = acc;              // the only writer of this element: no zero-fill, no reduction
      } else {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(acc.x), "f"(acc.y), "f"(acc.z), "f"(acc.w)
                     : "memory");
      }
    }
    if (epi.kind != 0) tile_epilogue(epi, y, ldy, M, N, tile, split, splits, tid);
  } else {
    if (splits > 1) {
      cluster_arrive();
      cluster_wait();
    }
    const float* red = reinterpret_cast<const float*>(base);
    const uint32_t red_addr = smem_u32(red);
    const int R = kTileN / splits;                           // weight rows reduced by this CTA
    const int qpr = R / 4;                                   // float4 chunks per batch row
    for (int item = tid; item < MP * qpr; item += kThreads) {
      const int m = item / qpr, c = (item - m * qpr) * 4 + split * R;      // c: row offset inside the tile
      const int n = tile * kTileN + c;
      if (m >= M || n >= N) continue;
      const uint32_t off = red_addr + (uint32_t)(m * kTileN + c) * 4u;
      float4 p[8];
#pragma unroll
      for (int sidx = 0; sidx < 8; ++sidx) {                 // all remote loads in flight before the first add
        p[sidx] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (sidx < splits && (mode == 0 || sidx == split)) {
          uint32_t ra;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(off), "r"(sidx));
          asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(p[sidx].x), "=f"(p[sidx].y), "=f"(p[sidx].z), "=f"(p[sidx].w)
                       : "r"(ra)
                       : "memory");
        }
      }
      float4 acc = p[0];
#pragma unroll
      for (int sidx = 1; sidx < 8; ++sidx) {
        acc.x += p[sidx].x; acc.y += p[sidx].y; acc.z += p[sidx].z; acc.w += p[sidx].w;
      }
      if (bias != nullptr) {
        const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + n));
        acc.x += bv.x; acc.y += bv.y; acc.z += bv.z; acc.w += bv.w;
      }
      float4* dst = reinterpret_cast<float4*>(y + (size_t)m * ldy + n);
      if (accumulate) {
        const float4 o = *dst;
        acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
      }
      *dst = acc;
    }
    if (splits > 1) {
      cluster_arrive();                                      // nobody leaves while a peer may still read its tile
      cluster_wait();
    }
  }
  if (tid == 0) STAMP(7);
  tc_fence_before();
  __syncthreads();
  if (tid == 32 && link.done_flag != nullptr) {              // every global access of this CTA precedes the barrier above
    asm volatile("fence.proxy.async;" ::: "memory");
    __threadfence();
    atomicAdd(link.done_flag, 1u);
  }
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(MP) : "memory");
  }
  if (tid == 0) {
    STAMP(8);
    if (dbg && blockIdx.x == 0 && blockIdx.y == 0) g_stamps[(slot % kStampSlots) * 12 + 11] = gtimer();
  }
}

// fp32 [N,K] -> bf16 hi / lo [N,K] and (optionally) transposed hi / lo [K,N]
__global__ void split_bf16_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                  __nv_bfloat16* __restrict__ hi_t, __nv_bfloat16* __restrict__ lo_t, int N, int K) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;               // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int n = n0 + j, k = k0 + tx;
    float v = 0.f;
    if (n < N && k < K) {
      v = w[(size_t)n * K + k];
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      hi[(size_t)n * K + k] = h;
      lo[(size_t)n * K + k] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
    tile[j][tx] = v;
  }
  if (hi_t == nullptr) return;
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int k = k0 + j, n = n0 + tx;
    if (n < N && k < K) {
      const float v = tile[tx][j];
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      hi_t[(size_t)k * N + n] = h;
      lo_t[(size_t)k * N + n] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
}

// split-K merge strategy (read once): VLN_GEMM_VARIANT = red4 (default: vector reductions into y) |
// c8 | c4 | c2 | c1 (splits of a tile form a cluster and reduce through DSMEM; measured slower)
struct Variant {
  int cap = 8, mode = 2, dbg = 0, bulk = 0;
  Variant() {
    dbg = getenv("VLN_GEMM_STAMPS") != nullptr;
    // VLN_GEMM_BULK=1: split-K merge through cp.reduce.async.bulk rows instead of REDG.128 (measured slower in the
    // iteration: 4.45 vs 4.36 ms, the wait for the bulk group's completion outweighs the shorter issue phase)
    bulk = getenv("VLN_GEMM_BULK") && getenv("VLN_GEMM_BULK")[0] == '1';
    const char* e = getenv("VLN_GEMM_VARIANT");
    if (!e) return;
    if (!strcmp(e, "c8")) mode = 0;
    else if (!strcmp(e, "c4")) mode = 0, cap = 4;
    else if (!strcmp(e, "c2")) mode = 0, cap = 2;
    else if (!strcmp(e, "c1")) mode = 0, cap = 1;
  }
};
const Variant& variant() {
  static Variant v;
  return v;
}

struct Second {                                            // second problem of a paired launch (same N, K, M, ld's)
  const void* w_hi = nullptr;
  const void* w_lo = nullptr;
  const float* x = nullptr;
  float* y = nullptr;
};

template <int MP, int ST>
int launch_linear_st(const void* w_hi, const void* w_lo, int N, int K, const float* x, int ldx, int M, const float* bias,
                     float* y, int ldy, int splits, int accumulate, cudaStream_t stream, Second sec, int m_tiles, Epi epi) {
  CUtensorMap tm_hi, tm_lo;
  int rc = vln_make_tmap_2d(&tm_hi, w_hi, (uint64_t)N, (uint64_t)K, (uint64_t)K, kBK, kTileN, 1);
  if (rc) return rc;
  rc = vln_make_tmap_2d(&tm_lo, w_lo, (uint64_t)N, (uint64_t)K, (uint64_t)K, kBK, kTileN, 1);
  if (rc) return rc;
  CUtensorMap tm_x;                                         // rows >= M are out of bounds: the TMA unit zero-fills them
  rc = vln_make_tmap_2d_f32(&tm_x, x, (uint64_t)M, (uint64_t)K, (uint64_t)ldx, kBK, MP);
  if (rc) return rc;
  CUtensorMap tm_hi1 = tm_hi, tm_lo1 = tm_lo, tm_x1 = tm_x;
  if (sec.y) {
    if ((rc = vln_make_tmap_2d(&tm_hi1, sec.w_hi, (uint64_t)N, (uint64_t)K, (uint64_t)K, kBK, kTileN, 1))) return rc;
    if ((rc = vln_make_tmap_2d(&tm_lo1, sec.w_lo, (uint64_t)N, (uint64_t)K, (uint64_t)K, kBK, kTileN, 1))) return rc;
    if ((rc = vln_make_tmap_2d_f32(&tm_x1, sec.x, (uint64_t)M, (uint64_t)K, (uint64_t)ldx, kBK, MP))) return rc;
  }
  static bool configured = false;
  if (!configured) {
    VLN_CHECK_CUDA(cudaFuncSetAttribute(linear_bf16x3_kernel<MP, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<MP, ST>::kSmem));
    configured = true;
  }
  const int tiles = (N + kTileN - 1) / kTileN;
  const int mode = variant().mode;
  if (mode == 2 && !accumulate && m_tiles == 1) VLN_CHECK_CUDA(cudaMemset2DAsync(y, (size_t)ldy * 4, 0, (size_t)N * 4, M, stream));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(tiles, m_tiles > 1 ? m_tiles : splits, sec.y ? 2 : 1);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = Cfg<MP, ST>::kSmem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (mode != 2 && splits > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;      // the splits of one weight tile = one cluster
    attr[na].val.clusterDim.x = 1;
    attr[na].val.clusterDim.y = splits;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (accumulate && vln_pdl_enabled()) {                    // part of a kernel chain (no memset node in front of it)
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  if (!(accumulate && vln_pdl_enabled())) vln_chain_break(stream);   // behind a memset / plain stream order: nothing to poll
  const ChainLink link = vln_chain_link(stream, cfg.gridDim.x * cfg.gridDim.y * cfg.gridDim.z);
  VLN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, linear_bf16x3_kernel<MP, ST>, tm_hi, tm_lo, tm_x, tm_hi1, tm_lo1, tm_x1, M, N, K, bias, y,
                                    sec.y ? sec.y : y, ldy, splits,
                                    accumulate, (m_tiles > 1 || mode == 2) ? (variant().bulk ? 3 : 2) : mode, variant().dbg, m_tiles,
                                    epi, link));
  return 0;
}

// ring depth by the k-blocks a CTA runs (see Cfg); VLN_GEMM_STAGES=0 always takes the deepest ring
template <int MP>
int launch_linear(const void* w_hi, const void* w_lo, int N, int K, const float* x, int ldx, int M, const float* bias,
                  float* y, int ldy, int splits, int accumulate, cudaStream_t stream, Second sec = Second(), int m_tiles = 1,
                  Epi epi = Epi()) {
  static const bool adapt = !(getenv("VLN_GEMM_STAGES") && getenv("VLN_GEMM_STAGES")[0] == '0');
  const int nkb = K / kBK;
  const int n_iter = m_tiles > 1 ? nkb : (nkb + splits - 1) / splits;
  constexpr int kDeep = MP == 64 ? 3 : 2;
  if (adapt && n_iter == 1)
    return launch_linear_st<MP, 1>(w_hi, w_lo, N, K, x, ldx, M, bias, y, ldy, splits, accumulate, stream, sec, m_tiles, epi);
  if (adapt && n_iter == 2)
    return launch_linear_st<MP, 2>(w_hi, w_lo, N, K, x, ldx, M, bias, y, ldy, splits, accumulate, stream, sec, m_tiles, epi);
  return launch_linear_st<MP, kDeep>(w_hi, w_lo, N, K, x, ldx, M, bias, y, ldy, splits, accumulate, stream, sec, m_tiles, epi);
}

}  // namespace

extern "C" int vln_linear_bf16x3(const void* w_hi, const void* w_lo, int N, int K, const float* x, int ldx, int M,
                                 const float* bias, float* y, int ldy, int accumulate, int splits, void* stream) {
  VLN_REQUIRE(w_hi && w_lo && x && y && N > 0 && M > 0, "bad arguments");
  VLN_REQUIRE(K > 0 && K % kBK == 0, "K must be a positive multiple of 64");
  VLN_REQUIRE(M <= 128, "at most 128 activation rows");
  VLN_REQUIRE(((uintptr_t)x & 15) == 0 && ldx % 4 == 0, "x rows must be 16-byte aligned");
  VLN_REQUIRE(N % 4 == 0 && ((uintptr_t)y & 15) == 0 && ldy % 4 == 0 && (!bias || ((uintptr_t)bias & 15) == 0),
              "N must be a multiple of 4 and y / bias 16-byte aligned");
  VLN_REQUIRE(((uintptr_t)w_hi & 15) == 0 && ((uintptr_t)w_lo & 15) == 0, "weights must be 16-byte aligned");
  const int nkb = K / kBK;
  const int tiles = (N + kTileN - 1) / kTileN;
  if (splits <= 0) splits = 148 / tiles;                       // one CTA per SM, one wave
  if (splits > nkb) splits = nkb;
  int s = 1;                                                   // cluster size: power of two, at most 8 (portable limit)
  while (s * 2 <= splits && s * 2 <= variant().cap) s *= 2;
  if (variant().mode == 2) s = splits;
  if (M <= 64) return launch_linear<64>(w_hi, w_lo, N, K, x, ldx, M, bias, y, ldy, s, accumulate, (cudaStream_t)stream);
  return launch_linear<128>(w_hi, w_lo, N, K, x, ldx, M, bias, y, ldy, s, accumulate, (cudaStream_t)stream);
}

// y += x W^T with a tile epilogue (see Epi): the shapes of the decoder step's H-wide products only (N = H outputs).
static int linear_epi(const void* w_hi, const void* w_lo, int N, int K, const float* x, int ldx, int M, float* y, int ldy,
                      const Epi& epi, void* stream) {
  VLN_REQUIRE(w_hi && w_lo && x && y && N > 0 && M > 0 && epi.ctr, "bad arguments");
  VLN_REQUIRE(K > 0 && K % kBK == 0, "K must be a positive multiple of 64");
  VLN_REQUIRE(M <= 128, "at most 128 activation rows");
  VLN_REQUIRE(N % 8 == 0 && N <= 8 * kTileN, "epilogue shapes: N a multiple of 8, at most 8 weight tiles");
  VLN_REQUIRE(((uintptr_t)x & 15) == 0 && ldx % 4 == 0 && ((uintptr_t)y & 15) == 0 && ldy % 4 == 0, "x / y rows must be 16-byte aligned");
  VLN_REQUIRE(((uintptr_t)w_hi & 15) == 0 && ((uintptr_t)w_lo & 15) == 0, "weights must be 16-byte aligned");
  VLN_REQUIRE(epi.p >= 0.f && epi.p < 1.f && (epi.p == 0.f || epi.rng), "dropout needs 0 <= p < 1 and an rng state");
  VLN_REQUIRE(variant().mode == 2, "tile epilogues need the vector-reduction split-K merge");
  const int nkb = K / kBK;
  const int tiles = (N + kTileN - 1) / kTileN;
  int splits = 148 / tiles;                                    // one CTA per SM, one wave: every CTA of a tile is resident
  if (splits > nkb) splits = nkb;
  if (M <= 64) return launch_linear<64>(w_hi, w_lo, N, K, x, ldx, M, nullptr, y, ldy, splits, 1, (cudaStream_t)stream, Second(), 1, epi);
  return launch_linear<128>(w_hi, w_lo, N, K, x, ldx, M, nullptr, y, ldy, splits, 1, (cudaStream_t)stream, Second(), 1, epi);
}

extern "C" int vln_linear_state_fwd(const void* w_hi, const void* w_lo, int N, int K, const float* x, int ldx, int M, float* y,
                                    int ldy, float* xh_next, int ld_xh, float* hq_next, float* hc_cur, float p,
                                    const uint64_t* rng, uint64_t off_q, uint64_t off_c, unsigned int* counters, void* stream) {
  VLN_REQUIRE(!xh_next || (ld_xh % 4 == 0 && ((uintptr_t)xh_next & 15) == 0), "xh rows must be 16-byte aligned");
  Epi e = Epi();
  e.kind = 1; e.ctr = counters; e.p = p; e.rng = rng; e.off_q = off_q; e.off_c = off_c;
  e.xh_next = xh_next; e.ld_xh = ld_xh; e.hq_next = hq_next; e.hc_cur = hc_cur;
  return linear_epi(w_hi, w_lo, N, K, x, ldx, M, y, ldy, e, stream);
}

extern "C" int vln_linear_state_bwd(const void* w_hi, const void* w_lo, int N, int K, const float* x, int ldx, int M, float* y,
                                    int ldy, const float* d_hc, const float* d_xh_next, int ld_dxh, const float* htilde,
                                    int ld_h, float* d_src, float p, const uint64_t* rng, uint64_t off_q, uint64_t off_c,
                                    unsigned int* counters, void* stream) {
  VLN_REQUIRE(d_hc && d_xh_next && htilde && d_src, "bad arguments");
  VLN_REQUIRE(ld_dxh % 4 == 0 && ld_h % 4 == 0 && (((uintptr_t)d_hc | (uintptr_t)d_xh_next | (uintptr_t)htilde | (uintptr_t)d_src) & 15) == 0,
              "operand rows must be 16-byte aligned");
  Epi e = Epi();
  e.kind = 2; e.ctr = counters; e.p = p; e.rng = rng; e.off_q = off_q; e.off_c = off_c;
  e.d_hc = d_hc; e.d_xh_next = d_xh_next; e.ld_dxh = ld_dxh; e.htilde = htilde; e.ld_h = ld_h; e.d_src = d_src;
  return linear_epi(w_hi, w_lo, N, K, x, ldx, M, y, ldy, e, stream);
}

// Tall activations (the encoder's input projection over all B*L token rows, units.py:58-63; the critic over all
// T*B states, policy.py:263): blocks of 128 activation rows x 128 weight rows per CTA, whole K, plain stores.
extern "C" int vln_linear_bf16x3_tall(const void* w_hi, const void* w_lo, int N, int K, const float* x, int ldx, int M,
                                      const float* bias, float* y, int ldy, int accumulate, void* stream) {
  VLN_REQUIRE(w_hi && w_lo && x && y && N > 0 && M > 0, "bad arguments");
  VLN_REQUIRE(K > 0 && K % kBK == 0, "K must be a positive multiple of 64");
  VLN_REQUIRE(((uintptr_t)x & 15) == 0 && ldx % 4 == 0, "x rows must be 16-byte aligned");
  VLN_REQUIRE(N % 4 == 0 && ((uintptr_t)y & 15) == 0 && ldy % 4 == 0 && (!bias || ((uintptr_t)bias & 15) == 0),
              "N must be a multiple of 4 and y / bias 16-byte aligned");
  VLN_REQUIRE(((uintptr_t)w_hi & 15) == 0 && ((uintptr_t)w_lo & 15) == 0, "weights must be 16-byte aligned");
  const int m_tiles = (M + 127) / 128;
  VLN_REQUIRE(m_tiles <= 65535, "too many activation rows");
  if (m_tiles == 1) return vln_linear_bf16x3(w_hi, w_lo, N, K, x, ldx, M, bias, y, ldy, accumulate, 0, stream);
  return launch_linear<128>(w_hi, w_lo, N, K, x, ldx, M, bias, y, ldy, 1, accumulate, (cudaStream_t)stream, Second(), m_tiles);
}

// Two independent products of identical shape in one launch: y0 += x0 W0^T, y1 += x1 W1^T (both accumulate).
extern "C" int vln_linear_bf16x3_pair(const void* w0_hi, const void* w0_lo, const float* x0, float* y0, const void* w1_hi,
                                      const void* w1_lo, const float* x1, float* y1, int N, int K, int ldx, int M, int ldy,
                                      void* stream) {
  VLN_REQUIRE(w0_hi && w0_lo && x0 && y0 && w1_hi && w1_lo && x1 && y1 && N > 0 && M > 0, "bad arguments");
  VLN_REQUIRE(K > 0 && K % kBK == 0 && M <= 128 && N % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0, "bad shape");
  VLN_REQUIRE(((uintptr_t)x0 & 15) == 0 && ((uintptr_t)x1 & 15) == 0 && ((uintptr_t)y0 & 15) == 0 && ((uintptr_t)y1 & 15) == 0,
              "operands must be 16-byte aligned");
  const int nkb = K / kBK;
  const int tiles = (N + kTileN - 1) / kTileN;
  int splits = 148 / (2 * tiles);                              // both problems together fill one wave
  if (splits < 1) splits = 1;
  if (splits > nkb) splits = nkb;
  VLN_REQUIRE(variant().mode == 2, "paired launches need the vector-reduction split-K merge");
  Second sec;
  sec.w_hi = w1_hi; sec.w_lo = w1_lo; sec.x = x1; sec.y = y1;
  if (M <= 64) return launch_linear<64>(w0_hi, w0_lo, N, K, x0, ldx, M, nullptr, y0, ldy, splits, 1, (cudaStream_t)stream, sec);
  return launch_linear<128>(w0_hi, w0_lo, N, K, x0, ldx, M, nullptr, y0, ldy, splits, 1, (cudaStream_t)stream, sec);
}

extern "C" int vln_debug_gemm_stamps(unsigned long long* out_host /*[64*12]*/, unsigned int* n) {
  VLN_CHECK_CUDA(cudaDeviceSynchronize());
  VLN_CHECK_CUDA(cudaMemcpyFromSymbol(out_host, g_stamps, sizeof(unsigned long long) * kStampSlots * 12));
  VLN_CHECK_CUDA(cudaMemcpyFromSymbol(n, g_stamp_n, sizeof(unsigned int)));
  return 0;
}

extern "C" int vln_split_bf16(const float* w, void* hi, void* lo, void* hi_t, void* lo_t, int N, int K, void* stream) {
  VLN_REQUIRE(w && hi && lo && N > 0 && K > 0, "bad arguments");
  VLN_REQUIRE((hi_t == nullptr) == (lo_t == nullptr), "transposed outputs come as a pair");
  split_bf16_kernel<<<dim3((K + 31) / 32, (N + 31) / 32), dim3(32, 8), 0, (cudaStream_t)stream>>>(
      w, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, (__nv_bfloat16*)hi_t, (__nv_bfloat16*)lo_t, N, K);
  VLN_LAUNCH_OK();
  return 0;
}
