// Persistent length-masked LSTM recurrence on the 5th-generation tensor cores (tcgen05), forward and
// backward-through-time: the packed nn.LSTM of EncoderLSTM (units.py:58-71), one launch per layer for both
// directions.  Same contract as csrc/lstm_seq.cu (the mma.sync version, kept behind VLN_LSTM_VARIANT=mma);
// what changes is where the recurrent product runs:
//
//   * a thread-block CLUSTER of C = H/32 CTAs owns NB (16, 24 or 32) batch rows for the whole sequence; CTA r owns
//     hidden units [32r, 32r+32), i.e. 128 gate rows (i,f,g,o x 32) of W_hh;
//   * those 128 x H weights stay resident in TENSOR MEMORY for all timesteps as the A operand of
//     tcgen05.mma (bf16 hi and lo halves, two bf16 per 32-bit column: lane = gate row, H/2 + H/2 columns),
//     written once with tcgen05.st.  A-from-TMEM matters at this shape: with A in shared memory every
//     128xNBx16 MMA would re-read 4 KB of weights (32 cycles at 128 B/clk) against an 8-cycle MMA;
//   * h_{t-1} is the B operand: bf16 hi / lo, K-major 128-byte-swizzled tiles in shared memory, double buffered.
//     Every CTA computes gates[128 x NB] = W_r . h_{t-1} with 3 x H/16 MMAs (hi.hi + hi.lo + lo.hi, fp32
//     accumulator in TMEM), four warps read the accumulator (tcgen05.ld), add the input projection, apply the
//     gate non-linearity; all eight warps do the cell update for (unit, row) pairs and send the new h slice
//     straight into every peer's NEXT B tile with 16-byte st.async stores that count on the receiver's mbarrier
//     (no cluster barrier per step);
//   * backward: A = W_hh^T restricted to the CTA's 128 gate rows (H x 128, H/128 M-tiles, resident in TMEM),
//     B = this step's dgates (written locally), D = partial dh_{t-1}[H x NB], reduce-scattered to the owning
//     CTAs with st.async and summed there in a fixed order (deterministic).
//
// NB = 16 keeps a B=64 bidirectional layer at 8 clusters: the mma.sync version (8 rows per cluster) needs 16
// clusters of 8 CTAs where only 15 are resident on a B200 — two waves, i.e. twice the 80-step latency chain.
// Roofline class: latency (L serial steps); tensor work per step and CTA is 3 x H/16 MMAs of 128 x NB x 16.
#include <cstdlib>
#include <cstring>

#include "common.cuh"

// optional phase stamps (VLN_LSTM_STAMPS=1): CTA (0,0) of the forward kernel, step 10 and the kernel's prologue / end
__device__ unsigned long long g_tc_stamps[16];
#define TSTAMP(i, cond)                                                                              \
  do {                                                                                               \
    if (dbg && blockIdx.x == 0 && blockIdx.y == 0 && (cond)) g_tc_stamps[i] = (unsigned long long)clock64(); \
  } while (0)

namespace {

constexpr int kThreads = 256;
constexpr int kHS = 32;         // hidden units per CTA
constexpr int kRows = 4 * kHS;  // gate rows per CTA
constexpr int kGP = kRows + 4;  // padded row of the activated-gate buffer

// Gate non-linearities on the SFU (ex2.approx + rcp.approx, ~2 ulp each): the precise expf / tanhf / IEEE division cost
// ~2 900 cycles per step for the two rows a thread updates, more than the tensor-core product itself.  Absolute error
// ~1e-7 per activation (tanh by 1 - 2 / (e^2x + 1): exact limits at +-inf, cancellation only below |x| ~ 1e-3).
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanhf_(float x) { return 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f); }
__device__ __forceinline__ void cluster_sync_all() {
  cluster_arrive();
  cluster_wait();
}
// mbarrier wait that traps instead of spinning forever (a protocol bug then surfaces as a launch failure, not a hung GPU)
__device__ __forceinline__ void mbar_wait_g(uint64_t* bar, uint32_t parity) {
  uint32_t n = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++n > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory operand: 8-row groups of 128-byte rows, SBO = 1024 B, descriptor version 1
__device__ __forceinline__ uint64_t sdesc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// D[tmem] (+)= A[tmem] . B[smem]   (A: lane = row, two bf16 per column; B: K-major swizzled tile).  Called by a whole
// converged warp with warp-uniform operands; one elected lane issues.
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]  (both K-major swizzled tiles): used for the lo half of the weights when H = 512
// leaves no tensor-memory columns for it
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
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
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
// NB (16, 24 or 32) consecutive accumulator columns of this thread's TMEM lane
template <int NB>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t* v) {
#pragma unroll
  for (int cg = 0; cg < NB / 16; ++cg) tmem_ld16(taddr + cg * 16, v + cg * 16);
  if constexpr (NB % 16 == 8) tmem_ld8(taddr + (NB / 16) * 16, v + (NB / 16) * 16);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t bf16_bits(float v) { return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v)); }
__device__ __forceinline__ float bf16_val(uint32_t bits) { return __uint_as_float(bits << 16); }
// two fp32 values -> packed bf16 pairs (first value in the low half): hi and the residual lo
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  const uint32_t ah = bf16_bits(a), bh = bf16_bits(b);
  hi = ah | (bh << 16);
  lo = bf16_bits(a - bf16_val(ah)) | (bf16_bits(b - bf16_val(bh)) << 16);
}
// 16-byte asynchronous DSMEM store into CTA `rank`, counted (16 bytes) on that CTA's mbarrier
__device__ __forceinline__ void dsmem_st_async_v4(void* local_ptr, uint64_t* local_bar, uint32_t rank, uint4 v) {
  uint32_t a = smem_u32(local_ptr), m = smem_u32(local_bar), ra, rm;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rm) : "r"(m), "r"(rank));
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(ra),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(rm)
               : "memory");
}

struct DirF {
  const float* xproj;   // [B,L,4H]
  const float* w_hh;    // [4H,H]
  float* out;           // [B,L,ld_out] + column offset of this direction
  float* acts;          // [B,L,4H]
  float* cs;            // [B,L,H]
  float* h_last;        // [B,ld_last] + column offset
  float* c_last;
  int reverse;
};
struct DirB {
  const float* w_hh;
  const float* acts;
  const float* cs;
  const float* d_out;   // [B,L,ld_out] + column offset (may be NULL)
  const float* d_hlast; // [B,ld_last] + column offset (may be NULL)
  const float* d_clast;
  float* d_xproj;       // [B,L,4H], pre-zeroed
  int reverse;
};

// instruction descriptor: D = F32, A = B = BF16, both K-major, N = NB, M = 128
template <int NB>
__device__ __forceinline__ constexpr uint32_t make_idesc() {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------
template <int H, int NB>
struct LayF {
  static constexpr int KB = H / 64;                       // 64-wide k-blocks of the B operand (h_{t-1})
  static constexpr int kTile = NB * 128;                  // bytes of one k-block tile: NB rows x 128 B
  static constexpr int kHBytes = 2 * 2 * KB * kTile;      // [buffer][hi/lo][k-block]
  // H = 512: hi (256 columns) + lo (256) + accumulator exceed the 512 tensor-memory columns, so the lo half of the
  // weights lives in shared memory as a K-major 128B-swizzled A operand (KB tiles of 128 rows x 128 B)
  static constexpr bool kLoSmem = H > 256;
  static constexpr int kALoOff = kHBytes;
  static constexpr int kALoBytes = kLoSmem ? KB * kRows * 128 : 0;
  static constexpr int kGOff = kHBytes + kALoBytes;       // float g[NB][kGP]: activated gates
  static constexpr int kStageOff = kGOff + NB * kGP * 4;  // uint16 stage[hi/lo][NB][32]: this CTA's new h slice
  static constexpr int kLenOff = kStageOff + 2 * NB * kHS * 2;
  static constexpr int kBarOff = kLenOff + NB * 4;
  static constexpr int kSmem = kBarOff + 64 + 1024;       // + barriers/TMEM slot + 1 KB alignment slack
  static constexpr int kColD = 0, kColAhi = 32, kColAlo = 32 + H / 2;
  static constexpr int kTmemCols = (32 + H) <= 256 ? 256 : 512;
  static_assert(kLoSmem ? (32 + H / 2 <= 512) : (32 + H <= 512), "tensor-memory columns");
  static_assert(kSmem <= 227 * 1024, "shared memory");
};

template <int H, int NB>
__global__ void __launch_bounds__(kThreads, 1)
lstm_tc_fwd_kernel(DirF d0, DirF d1, const int32_t* __restrict__ lengths, int B, int L, int ld_out, int ld_last, int dbg) {
  using S = LayF<H, NB>;
  constexpr int C = H / kHS, KB = S::KB, KS = H / 16, RPT = NB / 8;
  const DirF d = blockIdx.y == 0 ? d0 : d1;
  const float* __restrict__ xproj = d.xproj;
  const float* __restrict__ w_hh = d.w_hh;
  float* __restrict__ out = d.out;
  float* __restrict__ acts = d.acts;
  float* __restrict__ cs = d.cs;
  const int reverse = d.reverse;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* hb = base;
  const uint32_t hb_addr = smem_u32(hb);
  float* g = reinterpret_cast<float*>(base + S::kGOff);
  uint16_t* stage = reinterpret_cast<uint16_t*>(base + S::kStageOff);
  int* s_len = reinterpret_cast<int*>(base + S::kLenOff);
  uint64_t* bar_h = reinterpret_cast<uint64_t*>(base + S::kBarOff);   // [2]: all hi/lo bytes of h buffer k have arrived
  uint64_t* bar_acc = bar_h + 2;                                       // this step's MMAs have completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_h + 3);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform by construction: lets ptxas keep MMA operands in uniform registers
  const int rank = (int)cluster_ctarank();
  const int b0 = (blockIdx.x / C) * NB;

  TSTAMP(8, tid == 0);
  for (int i = tid; i < S::kHBytes / 16; i += kThreads) reinterpret_cast<uint4*>(hb)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid < NB) s_len[tid] = (b0 + tid < B) ? min(lengths[b0 + tid], L) : 0;
  if (tid == 0) {
    mbar_init(&bar_h[0], 1);
    mbar_init(&bar_h[1], 1);
    mbar_init(bar_acc, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(S::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  int gmax = 0;
#pragma unroll
  for (int i = 0; i < NB; ++i) gmax = max(gmax, s_len[i]);
  gmax = __shfl_sync(0xffffffffu, gmax, 0);

  // resident weights: TMEM lane lr = gate*32 + unit  <->  global row gate*H + rank*32 + unit; K along the columns
  if (warp < 4) {
    const int lr = tid;
    const float* wrow = w_hh + (size_t)((lr >> 5) * H + rank * kHS + (lr & 31)) * H;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 2
    for (int j = 0; j < KS; ++j) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int qd = 0; qd < 4; ++qd) {
        const float4 f = __ldg(reinterpret_cast<const float4*>(wrow + j * 16) + qd);
        split_pair(f.x, f.y, hi[2 * qd], lo[2 * qd]);
        split_pair(f.z, f.w, hi[2 * qd + 1], lo[2 * qd + 1]);
      }
      tmem_st8(lane_base + S::kColAhi + 8 * j, hi);
      if constexpr (S::kLoSmem) {     // 16 k = two 16-byte chunks of row lr in k-block j / 4
        uint8_t* row = base + S::kALoOff + (size_t)(j >> 2) * (kRows * 128) + lr * 128;
        *reinterpret_cast<uint4*>(row + ((((j & 3) * 2 + 0) ^ (lr & 7)) << 4)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        *reinterpret_cast<uint4*>(row + ((((j & 3) * 2 + 1) ^ (lr & 7)) << 4)) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
      } else {
        tmem_st8(lane_base + S::kColAlo + 8 * j, lo);
      }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();              // the zero-filled h tiles (and the lo weights) are read through the async proxy
  tc_fence_before();
  TSTAMP(9, tid == 0);
  cluster_sync_all();               // barriers initialised + buffers zeroed cluster-wide before any DSMEM traffic
  tc_fence_after();
  TSTAMP(10, tid == 0);

  const int q = warp & 3;           // warps 4..7 move TMEM lane quarter q (= gate q) to shared memory
  const int ug = rank * kHS + lane; // global hidden unit of the cell-update thread (rows warp, warp + 8, ...)
  float xp[RPT][4];
  auto load_x = [&](int t) {
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const int n = warp + 8 * j;
      const bool ok = t >= 0 && t < s_len[n];
      const float* p = xproj + ((size_t)(b0 + n) * L + (ok ? t : 0)) * (4 * H) + ug;
#pragma unroll
      for (int k = 0; k < 4; ++k) xp[j][k] = ok ? __ldg(p + k * H) : 0.f;
    }
  };
  float c_reg[RPT], h_reg[RPT];
#pragma unroll
  for (int j = 0; j < RPT; ++j) c_reg[j] = h_reg[j] = 0.f;

  int t = reverse ? gmax - 1 : 0;
  const int dt = reverse ? -1 : 1;
  load_x(t);
  constexpr uint32_t idesc = make_idesc<NB>();
  for (int s = 0; s < gmax; ++s, t += dt) {
    const int cur = s & 1, nxt = cur ^ 1;
    const bool send = s + 1 < gmax;                      // the last step's h has no consumer
    TSTAMP(0, tid == 0 && s == 10);
    if (tid == 0 && send) mbar_expect_tx(&bar_h[nxt], NB * H * 4);
    if (warp == 0) {
      if (s > 0) mbar_wait_g(&bar_h[cur], (uint32_t)((s - 1) >> 1) & 1u);
      TSTAMP(1, tid == 0 && s == 10);
      tc_fence_after();
      const uint32_t bh = hb_addr + (uint32_t)((cur * 2 + 0) * KB * S::kTile);
      const uint32_t bl = hb_addr + (uint32_t)((cur * 2 + 1) * KB * S::kTile);
      const uint64_t dh0 = sdesc_sw128(bh), dl0 = sdesc_sw128(bl);
      [[maybe_unused]] const uint64_t alo0 = sdesc_sw128(hb_addr + S::kALoOff);
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const uint64_t inc = (uint64_t)(((ks >> 2) * S::kTile + (ks & 3) * 32) >> 4);   // address field is in 16-byte units
        umma_ts(tmem + S::kColD, tmem + S::kColAhi + 8 * ks, dh0 + inc, idesc, ks != 0);
        umma_ts(tmem + S::kColD, tmem + S::kColAhi + 8 * ks, dl0 + inc, idesc, 1u);
        if constexpr (S::kLoSmem) {
          const uint64_t ainc = (uint64_t)(((ks >> 2) * (kRows * 128) + (ks & 3) * 32) >> 4);
          umma_ss(tmem + S::kColD, alo0 + ainc, dh0 + inc, idesc, 1u);
        } else {
          umma_ts(tmem + S::kColD, tmem + S::kColAlo + 8 * ks, dh0 + inc, idesc, 1u);
        }
      }
      umma_commit(bar_acc);
      __syncwarp();
      TSTAMP(2, tid == 0 && s == 10);
    } else if (warp >= 4) {
      // W_r . h_{t-1} of my gate row (gate q, unit lane) for the NB batch rows: TMEM -> shared memory
      mbar_wait_g(bar_acc, (uint32_t)s & 1u);
      TSTAMP(3, tid == 128 && s == 10);
      tc_fence_after();
      uint32_t v[NB];
      tmem_ld_cols<NB>(tmem + ((uint32_t)(q * 32) << 16) + S::kColD, v);
#pragma unroll
      for (int n = 0; n < NB; ++n) g[n * kGP + q * kHS + lane] = __uint_as_float(v[n]);
      tc_fence_before();
      TSTAMP(4, tid == 128 && s == 10);
    }
    __syncthreads();
    TSTAMP(5, tid == 0 && s == 10);
    // gates + cell update for (row n = warp + 8j, unit lane); new h as bf16 hi / lo into the staging tile.  What the
    // peers' next product waits for — the h slices — leaves first; the global stores of this step (out, cs, activated
    // gates for the backward pass) and the prefetch of the next input projection follow, off the inter-CTA chain.
    float sv_i[RPT], sv_f[RPT], sv_g[RPT], sv_o[RPT];
    bool live[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const int n = warp + 8 * j;
      const float* gp = g + n * kGP + lane;
      const float ig = sigmoidf_(gp[0] + xp[j][0]), fg = sigmoidf_(gp[32] + xp[j][1]), gg = tanhf_(gp[64] + xp[j][2]),
                  og = sigmoidf_(gp[96] + xp[j][3]);
      const float c_new = fg * c_reg[j] + ig * gg;
      const float h_new = og * tanhf_(c_new);
      live[j] = t < s_len[n];                              // rows past their length keep their state, outputs stay zero
      if (live[j]) {
        c_reg[j] = c_new;
        h_reg[j] = h_new;
      }
      sv_i[j] = ig; sv_f[j] = fg; sv_g[j] = gg; sv_o[j] = og;
      const uint32_t hi16 = bf16_bits(h_reg[j]);
      stage[(0 * NB + n) * kHS + lane] = (uint16_t)hi16;
      stage[(1 * NB + n) * kHS + lane] = (uint16_t)bf16_bits(h_reg[j] - bf16_val(hi16));
      __syncwarp();
      // all-gather: the 32 units of row n are 4 hi + 4 lo chunks of 16 bytes, each goes to all C CTAs' next B tile
      // (issued per row, so the stores of one row travel while the next row is computed)
      if (send) {
#pragma unroll
        for (int it = 0; it < C / 4; ++it) {
          const int i = lane + 32 * it;
          const int peer = i % C, ch = i / C;              // ch 0..7: hi chunks 0..3, lo chunks 0..3
          const int hilo = ch >> 2, chunk = ch & 3;
          const uint4 val = *reinterpret_cast<const uint4*>(stage + (hilo * NB + n) * kHS + chunk * 8);
          const int k = rank * kHS + chunk * 8;            // first hidden unit of the chunk
          uint8_t* dst = hb + (size_t)((nxt * 2 + hilo) * KB + (k >> 6)) * S::kTile + n * 128 +
                         ((((k & 63) >> 3) ^ (n & 7)) << 4);
          dsmem_st_async_v4(dst, &bar_h[nxt], (uint32_t)peer, val);
        }
      }
    }
    TSTAMP(6, tid == 0 && s == 10);
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      if (live[j]) {
        const size_t o = (size_t)(b0 + warp + 8 * j) * L + t;
        out[o * ld_out + ug] = h_reg[j];
        cs[o * H + ug] = c_reg[j];
        float* ap = acts + o * (4 * H) + ug;
        ap[0] = sv_i[j]; ap[H] = sv_f[j]; ap[2 * H] = sv_g[j]; ap[3 * H] = sv_o[j];
      }
    }
    load_x(t + dt);                                        // next step's input projection (consumed after the next product)
    TSTAMP(7, tid == 0 && s == 10);
    // Buffer reuse needs no further barrier: a peer writes h[cur] again only in step s+1, which it enters after
    // it received THIS CTA's slice of step s — sent after this CTA's MMAs of step s (which read h[cur]) completed.
  }
  TSTAMP(11, tid == 0);
#pragma unroll
  for (int j = 0; j < RPT; ++j) {
    const int b = b0 + warp + 8 * j;
    if (b < B) {
      d.h_last[(size_t)b * ld_last + ug] = h_reg[j];
      d.c_last[(size_t)b * ld_last + ug] = c_reg[j];
    }
  }
  tc_fence_before();
  cluster_sync_all();               // no CTA exits while a peer could still address its shared memory
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(S::kTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------
// backward through time
// ------------------------------------------------------------------------------------------------------------
template <int H, int NB>
struct LayB {
  static constexpr int C = H / kHS;
  static constexpr int MT = H / 128;                      // 128-column M-tiles of dh
  static constexpr int kTile = NB * 128;                  // one 64-wide k-block of the dgates operand
  static constexpr int kDgBytes = 2 * 2 * kTile;          // [hi/lo][k-block]
  static constexpr bool kLoSmem = H > 256;                // lo half of W^T in shared memory: [MT][2 k-blocks][128 x 128 B]
  static constexpr int kALoOff = (kDgBytes + 1023) / 1024 * 1024;
  static constexpr int kALoBytes = kLoSmem ? MT * 2 * kRows * 128 : 0;
  static constexpr int kRecvOff = kALoOff + kALoBytes;    // float recv[2][C][NB/4][32][4]: partial dh for my units
  static constexpr int kRecvBytes = 2 * C * NB * kHS * 4;
  static constexpr int kLenOff = kRecvOff + kRecvBytes;
  static constexpr int kBarOff = kLenOff + NB * 4;
  static constexpr int kSmem = kBarOff + 64 + 1024;
  static constexpr int kColAhi = MT * 32 > 64 ? MT * 32 : 64, kColAlo = kColAhi + MT * 64;    // D[mt] at column mt * 32
  static constexpr int kColD(int mt) { return mt * 32; }
  static constexpr int kTmemCols = (MT * 32 + 2 * MT * 64) <= 256 ? 256 : 512;
  static_assert(kLoSmem ? (MT * 32 + MT * 64 <= 512) : (MT <= 2), "tensor-memory columns");
  static_assert(kSmem <= 227 * 1024, "shared memory");
};

template <int H, int NB>
__global__ void __launch_bounds__(kThreads, 1)
lstm_tc_bwd_kernel(DirB d0, DirB d1, const int32_t* __restrict__ lengths, int B, int L, int ld_out, int ld_last, int dbg) {
  using S = LayB<H, NB>;
  constexpr int C = S::C, MT = S::MT, RPT = NB / 8, KS = kRows / 16;
  const DirB d = blockIdx.y == 0 ? d0 : d1;
  const float* __restrict__ w_hh = d.w_hh;
  const float* __restrict__ acts = d.acts;
  const float* __restrict__ cs = d.cs;
  const float* __restrict__ d_out = d.d_out;
  float* __restrict__ d_xproj = d.d_xproj;
  const int reverse = d.reverse;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* dgb = base;
  const uint32_t dg_addr = smem_u32(dgb);
  float* recv = reinterpret_cast<float*>(base + S::kRecvOff);
  int* s_len = reinterpret_cast<int*>(base + S::kLenOff);
  uint64_t* bar_r = reinterpret_cast<uint64_t*>(base + S::kBarOff);   // [2]: all partials of recv[k] have arrived
  uint64_t* bar_acc = bar_r + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_r + 3);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform by construction: lets ptxas keep MMA operands in uniform registers
  const int rank = (int)cluster_ctarank();
  const int b0 = (blockIdx.x / C) * NB;

  if (tid < NB) s_len[tid] = (b0 + tid < B) ? min(lengths[b0 + tid], L) : 0;
  if (tid == 0) {
    mbar_init(&bar_r[0], 1);
    mbar_init(&bar_r[1], 1);
    mbar_init(bar_acc, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(S::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  int gmax = 0;
#pragma unroll
  for (int i = 0; i < NB; ++i) gmax = max(gmax, s_len[i]);
  gmax = __shfl_sync(0xffffffffu, gmax, 0);

  // resident W^T restricted to this CTA's gate rows: dh[k][n] = sum_lr W[lr][k] dg[n][lr];  A[m = column k][kk = local row];
  // M-tile mt is written by warps 4mt .. 4mt+3 (TMEM lane = k % 128)
  for (int mt = warp >> 2; mt < MT; mt += 2) {
    const int k = mt * 128 + (warp & 3) * 32 + lane;
    const int row = (warp & 3) * 32 + lane;                 // TMEM lane = row of the A tile
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll 2
    for (int j = 0; j < KS; ++j) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const int lr0 = 16 * j + 2 * p, lr1 = lr0 + 1;
        const float w0 = __ldg(w_hh + (size_t)((lr0 >> 5) * H + rank * kHS + (lr0 & 31)) * H + k);
        const float w1 = __ldg(w_hh + (size_t)((lr1 >> 5) * H + rank * kHS + (lr1 & 31)) * H + k);
        split_pair(w0, w1, hi[p], lo[p]);
      }
      tmem_st8(lane_base + S::kColAhi + mt * 64 + 8 * j, hi);
      if constexpr (S::kLoSmem) {
        uint8_t* rp = base + S::kALoOff + (size_t)(mt * 2 + (j >> 2)) * (kRows * 128) + row * 128;
        *reinterpret_cast<uint4*>(rp + ((((j & 3) * 2 + 0) ^ (row & 7)) << 4)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        *reinterpret_cast<uint4*>(rp + ((((j & 3) * 2 + 1) ^ (row & 7)) << 4)) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
      } else {
        tmem_st8(lane_base + S::kColAlo + mt * 64 + 8 * j, lo);
      }
    }
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  fence_proxy_async();
  const int ug = rank * kHS + lane;
  float dh[RPT], dc[RPT];
#pragma unroll
  for (int j = 0; j < RPT; ++j) {
    const int b = b0 + warp + 8 * j;
    dh[j] = (b < B && d.d_hlast) ? d.d_hlast[(size_t)b * ld_last + ug] : 0.f;
    dc[j] = (b < B && d.d_clast) ? d.d_clast[(size_t)b * ld_last + ug] : 0.f;
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();

  // forward visited t = 0..gmax-1 (or gmax-1..0 when reversed); walk it backwards
  int t = reverse ? 0 : gmax - 1;
  const int dt = reverse ? 1 : -1;
  // saved activations / cell states / incoming gradient of the step are fetched one step ahead (registers), so the
  // L2 round trip overlaps the previous step's product and exchange instead of heading every step's critical path
  float pre[RPT][7];                                       // ig, fg, gg, og, c_t, c_{t-1}, d_out
  auto load_step = [&](int tt) {
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const int n = warp + 8 * j;
      const int len = s_len[n];
      const bool ok = tt >= 0 && tt < len;
      const size_t o = (size_t)(b0 + n) * L + (ok ? tt : 0);
      const float* ap = acts + o * (4 * H) + ug;
      const int tp = tt - (reverse ? -1 : 1);                        // the step that produced c_{prev}
      const bool okp = ok && tp >= 0 && tp < len;
#pragma unroll
      for (int k = 0; k < 4; ++k) pre[j][k] = ok ? __ldg(ap + k * H) : 0.f;
      pre[j][4] = ok ? __ldg(cs + o * H + ug) : 0.f;
      pre[j][5] = okp ? __ldg(cs + ((size_t)(b0 + n) * L + tp) * H + ug) : 0.f;
      pre[j][6] = (ok && d_out) ? __ldg(d_out + o * ld_out + ug) : 0.f;
    }
  };
  load_step(t);
  constexpr uint32_t idesc = make_idesc<NB>();
  for (int s = 0; s < gmax; ++s, t += dt) {
    const int buf = s & 1;
    const bool more = s + 1 < gmax;                      // dh of the step before the first one has no consumer
    if (tid == 0 && more) mbar_expect_tx(&bar_r[buf], C * NB * kHS * 4);
    float cur[RPT][7];
#pragma unroll
    for (int j = 0; j < RPT; ++j)
#pragma unroll
      for (int k = 0; k < 7; ++k) cur[j][k] = pre[j][k];
    load_step(t + dt);
    // pointwise gradient for (row n = warp + 8j, unit lane); dgates as bf16 hi / lo into the B tiles first (the
    // product waits for them), their global copy (d_xproj) after the product has been issued
    float dgs[RPT][4];
    bool live[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const int n = warp + 8 * j;
      float dg[4] = {0.f, 0.f, 0.f, 0.f};
      live[j] = t < s_len[n];
      if (live[j]) {
        const float ig = cur[j][0], fg = cur[j][1], gg = cur[j][2], og = cur[j][3];
        const float c1 = cur[j][4], c0 = cur[j][5];
        const float dht = dh[j] + cur[j][6];
        const float tc = tanhf_(c1);
        const float dct = dc[j] + dht * og * (1.f - tc * tc);
        dg[0] = dct * gg * ig * (1.f - ig);
        dg[1] = dct * c0 * fg * (1.f - fg);
        dg[2] = dct * ig * (1.f - gg * gg);
        dg[3] = dht * tc * og * (1.f - og);
        dc[j] = dct * fg;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) dgs[j][k] = dg[k];
      if (more) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int lr = k * kHS + lane;                             // local gate row = K index of the MMA
          const int kk = lr & 63;
          const uint32_t off = (uint32_t)((lr >> 6) * S::kTile + n * 128 + (((kk >> 3) ^ (n & 7)) << 4) + (kk & 7) * 2);
          const uint32_t hi = bf16_bits(dg[k]);
          *reinterpret_cast<uint16_t*>(dgb + off) = (uint16_t)hi;
          *reinterpret_cast<uint16_t*>(dgb + 2 * S::kTile + off) = (uint16_t)bf16_bits(dg[k] - bf16_val(hi));
        }
      }
    }
    auto store_dx = [&]() {
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        if (live[j]) {
          float* dp = d_xproj + ((size_t)(b0 + warp + 8 * j) * L + t) * (4 * H) + ug;
          dp[0] = dgs[j][0]; dp[H] = dgs[j][1]; dp[2 * H] = dgs[j][2]; dp[3 * H] = dgs[j][3];
        }
      }
    };
    if (!more) {
      store_dx();
      break;
    }
    fence_proxy_async();            // generic-proxy stores -> visible to the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      const uint64_t dh0 = sdesc_sw128(dg_addr), dl0 = sdesc_sw128(dg_addr + 2 * S::kTile);
      [[maybe_unused]] const uint64_t alo0 = sdesc_sw128(dg_addr + S::kALoOff);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const uint64_t inc = (uint64_t)(((ks >> 2) * S::kTile + (ks & 3) * 32) >> 4);
          umma_ts(tmem + mt * 32, tmem + S::kColAhi + mt * 64 + 8 * ks, dh0 + inc, idesc, ks != 0);
          umma_ts(tmem + mt * 32, tmem + S::kColAhi + mt * 64 + 8 * ks, dl0 + inc, idesc, 1u);
          if constexpr (S::kLoSmem) {
            const uint64_t ainc = (uint64_t)(((mt * 2 + (ks >> 2)) * (kRows * 128) + (ks & 3) * 32) >> 4);
            umma_ss(tmem + mt * 32, alo0 + ainc, dh0 + inc, idesc, 1u);
          } else {
            umma_ts(tmem + mt * 32, tmem + S::kColAlo + mt * 64 + 8 * ks, dh0 + inc, idesc, 1u);
          }
        }
      }
      umma_commit(bar_acc);
    }
    __syncwarp();
    store_dx();                       // under the tensor-core product
    {
      // reduce-scatter: TMEM lane = column k of M-tile mt -> owner CTA k / 32, slot [my rank][n / 4][k % 32][n % 4]
      const int qq = warp & 3;
      if ((warp >> 2) < MT) {
        mbar_wait_g(bar_acc, (uint32_t)s & 1u);
        tc_fence_after();
      }
      for (int mt = warp >> 2; mt < MT; mt += 2) {
        uint32_t v[NB];
        tmem_ld_cols<NB>(tmem + ((uint32_t)(qq * 32) << 16) + mt * 32, v);
        const uint32_t owner = (uint32_t)(mt * 4 + qq);
#pragma unroll
        for (int n4 = 0; n4 < NB / 4; ++n4) {
          float* dst = recv + ((size_t)((buf * C + rank) * (NB / 4) + n4) * kHS + lane) * 4;
          dsmem_st_async_v4(dst, &bar_r[buf], owner, make_uint4(v[4 * n4], v[4 * n4 + 1], v[4 * n4 + 2], v[4 * n4 + 3]));
        }
      }
      tc_fence_before();
    }
    mbar_wait_g(&bar_r[buf], (uint32_t)(s >> 1) & 1u);
    // owner: dh_{t-1}[n][my unit] = sum over the C partials in rank order (live rows); frozen rows pass dh through
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const int n = warp + 8 * j;
      if (t < s_len[n]) {
        float acc = 0.f;
#pragma unroll
        for (int r = 0; r < C; ++r) acc += recv[((size_t)((buf * C + r) * (NB / 4) + (n >> 2)) * kHS + lane) * 4 + (n & 3)];
        dh[j] = acc;
      }
    }
    // recv is double buffered: a peer's stores of step s+1 go to the other half, and its stores of step s+2 come
    // after it received this CTA's partials of step s+1, which are sent only after the reads above.  The dgates
    // tiles are rewritten in step s+1 only after this wait, i.e. after this CTA's MMAs of step s completed.
  }
  tc_fence_before();
  cluster_sync_all();               // no CTA exits while a peer could still address its shared memory
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(S::kTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------
template <typename K>
int max_active_clusters(K kernel, size_t smem, int C, int* out) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(C * 64, 1);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  VLN_CHECK_CUDA(cudaOccupancyMaxActiveClusters(out, kernel, &cfg));
  return 0;
}

template <typename K, typename D>
int launch_tc(K kernel, size_t smem, int C, int NB, int n_dir, int B, cudaStream_t stream, D d0, D d1,
              const int32_t* lengths, int L, int ld_out, int ld_last) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(((B + NB - 1) / NB) * C, n_dir);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  VLN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, d0, d1, lengths, B, L, ld_out, ld_last, (int)(getenv("VLN_LSTM_STAMPS") != nullptr)));
  return 0;
}

// The smallest NB in {16, 24, 32} whose launch fits in ONE wave of resident clusters (a second wave doubles the serial
// latency chain; a larger NB lengthens every step: MMA N, gate math and the h exchange all scale with it);
// VLN_LSTM_NB=16|24|32 forces it.
template <int H>
int pick_nb(int B, int n_dir, bool fwd) {
  static int max16[2] = {-1, -1};
  const char* e = getenv("VLN_LSTM_NB");
  if (e && !strcmp(e, "32")) return 32;
  if (e && !strcmp(e, "24")) return 24;
  if (e && !strcmp(e, "16")) return 16;
  int& m = max16[fwd ? 0 : 1];
  if (m < 0) {
    int v = 0;
    int rc;
    if (fwd) {
      cudaFuncSetAttribute(lstm_tc_fwd_kernel<H, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, LayF<H, 32>::kSmem);
      rc = max_active_clusters(lstm_tc_fwd_kernel<H, 32>, LayF<H, 32>::kSmem, H / kHS, &v);
    } else {
      cudaFuncSetAttribute(lstm_tc_bwd_kernel<H, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, LayB<H, 32>::kSmem);
      rc = max_active_clusters(lstm_tc_bwd_kernel<H, 32>, LayB<H, 32>::kSmem, H / kHS, &v);
    }
    m = (rc == 0 && v > 0) ? v : 14;
  }
  for (int nb = 16; nb < 32; nb += 8)
    if (((B + nb - 1) / nb) * n_dir <= m) return nb;
  return 32;
}

template <int H, int NB>
int launch_fwd(DirF d0, DirF d1, int n_dir, const int32_t* lengths, int B, int L, int ld_out, int ld_last, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    VLN_CHECK_CUDA(cudaFuncSetAttribute(lstm_tc_fwd_kernel<H, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, LayF<H, NB>::kSmem));
    if (H / kHS > 8)                                          // 16-CTA clusters (H = 512) are beyond the portable size
      VLN_CHECK_CUDA(cudaFuncSetAttribute(lstm_tc_fwd_kernel<H, NB>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    configured = true;
  }
  return launch_tc(lstm_tc_fwd_kernel<H, NB>, LayF<H, NB>::kSmem, H / kHS, NB, n_dir, B, stream, d0, d1, lengths, L, ld_out, ld_last);
}
template <int H, int NB>
int launch_bwd(DirB d0, DirB d1, int n_dir, const int32_t* lengths, int B, int L, int ld_out, int ld_last, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    VLN_CHECK_CUDA(cudaFuncSetAttribute(lstm_tc_bwd_kernel<H, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, LayB<H, NB>::kSmem));
    if (H / kHS > 8)
      VLN_CHECK_CUDA(cudaFuncSetAttribute(lstm_tc_bwd_kernel<H, NB>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    configured = true;
  }
  return launch_tc(lstm_tc_bwd_kernel<H, NB>, LayB<H, NB>::kSmem, H / kHS, NB, n_dir, B, stream, d0, d1, lengths, L, ld_out, ld_last);
}

}  // namespace

extern "C" int vln_debug_lstm_tc_stamps(unsigned long long* out_host /*[16]*/) {
  VLN_CHECK_CUDA(cudaDeviceSynchronize());
  VLN_CHECK_CUDA(cudaMemcpyFromSymbol(out_host, g_tc_stamps, sizeof(unsigned long long) * 16));
  return 0;
}

// Same contracts as vln_lstm_seq_fwd / vln_lstm_seq_bwd (csrc/lstm_seq.cu), which dispatch here by default.
int vln_lstm_tc_fwd(const float* const* xproj, const float* const* w_hh, const int32_t* lengths, float* out,
                    float* const* acts, float* const* cs, float* h_last, float* c_last, int B, int L, int H, int n_dir,
                    cudaStream_t stream) {
  DirF d[2] = {};
  for (int k = 0; k < n_dir; ++k)
    d[k] = DirF{xproj[k], w_hh[k], out + k * H, acts[k], cs[k], h_last + k * H, c_last + k * H, k};
  const int ld = n_dir * H;
  if (H == 256) {
    const int nb = pick_nb<256>(B, n_dir, true);
    if (nb == 16) return launch_fwd<256, 16>(d[0], d[1], n_dir, lengths, B, L, ld, ld, stream);
    if (nb == 24) return launch_fwd<256, 24>(d[0], d[1], n_dir, lengths, B, L, ld, ld, stream);
    return launch_fwd<256, 32>(d[0], d[1], n_dir, lengths, B, L, ld, ld, stream);
  }
  if (H == 128) {
    const int nb = pick_nb<128>(B, n_dir, true);
    if (nb == 16) return launch_fwd<128, 16>(d[0], d[1], n_dir, lengths, B, L, ld, ld, stream);
    if (nb == 24) return launch_fwd<128, 24>(d[0], d[1], n_dir, lengths, B, L, ld, ld, stream);
    return launch_fwd<128, 32>(d[0], d[1], n_dir, lengths, B, L, ld, ld, stream);
  }
  if (H == 512)     // Self-Monitor's uni-directional encoder: 16-CTA clusters, lo weights in shared memory, 16 rows per cluster
    return launch_fwd<512, 16>(d[0], d[1], n_dir, lengths, B, L, ld, ld, stream);
  vln_set_error("vln_lstm_seq_fwd: hidden size %d per direction is not supported (128, 256 or 512)", H);
  return -1;
}

int vln_lstm_tc_bwd(const float* const* w_hh, const int32_t* lengths, const float* const* acts, const float* const* cs,
                    const float* d_out, const float* d_hlast, const float* d_clast, float* const* d_xproj, int B, int L,
                    int H, int n_dir, cudaStream_t stream) {
  DirB d[2] = {};
  for (int k = 0; k < n_dir; ++k)
    d[k] = DirB{w_hh[k], acts[k], cs[k], d_out ? d_out + k * H : nullptr, d_hlast ? d_hlast + k * H : nullptr,
                d_clast ? d_clast + k * H : nullptr, d_xproj[k], k};
  const int ld = n_dir * H;
  if (H == 256) {
    const int nb = pick_nb<256>(B, n_dir, false);
    if (nb == 16) return launch_bwd<256, 16>(d[0], d[1], n_dir, lengths, B, L, ld, ld, stream);
    if (nb == 24) return launch_bwd<256, 24>(d[0], d[1], n_dir, lengths, B, L, ld, ld, stream);
    return launch_bwd<256, 32>(d[0], d[1], n_dir, lengths, B, L, ld, ld, stream);
  }
  if (H == 128) {
    const int nb = pick_nb<128>(B, n_dir, false);
    if (nb == 16) return launch_bwd<128, 16>(d[0], d[1], n_dir, lengths, B, L, ld, ld, stream);
    if (nb == 24) return launch_bwd<128, 24>(d[0], d[1], n_dir, lengths, B, L, ld, ld, stream);
    return launch_bwd<128, 32>(d[0], d[1], n_dir, lengths, B, L, ld, ld, stream);
  }
  if (H == 512) return launch_bwd<512, 16>(d[0], d[1], n_dir, lengths, B, L, ld, ld, stream);
  vln_set_error("vln_lstm_seq_bwd: hidden size %d per direction is not supported (128, 256 or 512)", H);
  return -1;
}
