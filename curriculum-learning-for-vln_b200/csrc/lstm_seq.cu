// Persistent length-masked LSTM recurrence (one direction of EncoderLSTM's nn.LSTM, units.py:58-71),
// forward and backward-through-time, one launch each.
//
// The reference packs the padded batch and calls cuDNN; here a THREAD-BLOCK CLUSTER owns a group
// of kNB batch rows for the whole sequence:
//   * CTA r of the cluster owns hidden units [32r, 32r+32): its 128 gate rows (i,f,g,o x 32) of
//     W_hh stay resident for all timesteps — in REGISTERS, as the A fragments of mma.m16n8k16 (bf16 hi
//     and lo halves of every weight: warp w holds rows 16w..16w+15, 128 registers per thread at H=256);
//   * every step each CTA computes its 128 gate pre-activations for the group's kNB = 8 rows (the MMA's
//     N) from the full h_{t-1}, kept in shared memory as bf16 hi/lo pairs: three tensor-core products
//     hi.hi + hi.lo + lo.hi with fp32 accumulation reproduce the fp32 matvec to ~2^-16 (the fp32 FMA
//     version of this loop was bound by LDS bandwidth at ~5 us per step); then the LSTM pointwise update
//     of its 32 units (c stays in registers), and its slice of h_t goes to every peer's shared memory
//     through asynchronous DSMEM stores that signal the receiver's mbarrier (no cluster barrier per step);
//   * rows stop at their own length (reverse rows start there) exactly like a packed sequence:
//     state frozen and output zero where t >= length;  steps past the group's longest row are skipped.
// Backward runs the same structure in reverse time: the pointwise gradient for the CTA's units,
// partial dh_{t-1} = dgates . W_hh over its 128 rows, reduce-scattered to the owning CTAs via DSMEM.
// dgates (= d xproj) goes to global memory; dW_hh / dW_ih / dx are single large GEMMs outside.
//
// Roofline class: latency (80 serial steps); per step and CTA 3 x 16 x H/16 MMAs (m16n8k16).
#include <cstdlib>
#include <cstring>

#include "common.cuh"

// optional phase stamps of one timestep (VLN_LSTM_STAMPS=1): CTA (0,0), thread 0, step 10 of the forward kernel
__device__ unsigned long long g_lstm_stamps[12];
#define LSTAMP0(i)                                                                                      \
  do {                                                                                                  \
    if (dbg && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) g_lstm_stamps[i] = (unsigned long long)clock64(); \
  } while (0)
#define LSTAMP(i)                                                                                       \
  do {                                                                                                  \
    if (dbg && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0 && s == 10) g_lstm_stamps[i] = (unsigned long long)clock64(); \
  } while (0)

namespace {

constexpr int kNB = 8;          // batch rows per cluster
constexpr int kHS = 32;         // hidden units per CTA
constexpr int kRows = 4 * kHS;  // gate rows per CTA
constexpr int kThreads = 256;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ void cluster_sync_all() {
  cluster_arrive();
  cluster_wait();
}
// Asynchronous DSMEM store that signals the DESTINATION CTA's mbarrier (4 bytes of its transaction
// count): the receiver waits on its own barrier for exactly the bytes it expects, so a timestep needs
// no cluster-wide barrier (barrier.cluster also drains every outstanding global store of the step and
// measured several microseconds per step on B200).
__device__ __forceinline__ void dsmem_st_async_f32(float* local_ptr, uint64_t* local_bar, uint32_t rank, float v) {
  uint32_t a = smem_u32(local_ptr), m = smem_u32(local_bar), ra, rm;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rm) : "r"(m), "r"(rank));
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(ra),
               "r"(__float_as_uint(v)), "r"(rm)
               : "memory");
}
// (default .acquire.cta semantics: the data arrives in THIS CTA's shared memory through the async proxy and is
// published by the barrier's transaction count, as with TMA; a cluster-scope acquire would add an L1 invalidate)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t bf16_bits(float v) { return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v)); }
__device__ __forceinline__ float bf16_val(uint32_t bits) { return __uint_as_float(bits << 16); }
// pack two fp32 values into bf16 pairs (first value in the low half): hi and the residual lo
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  const uint32_t ah = bf16_bits(a), bh = bf16_bits(b);
  hi = ah | (bh << 16);
  lo = bf16_bits(a - bf16_val(ah)) | (bf16_bits(b - bf16_val(bh)) << 16);
}
__device__ __forceinline__ void dsmem_st_async_u32(void* local_ptr, uint64_t* local_bar, uint32_t rank, uint32_t v) {
  uint32_t a = smem_u32(local_ptr), m = smem_u32(local_bar), ra, rm;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rm) : "r"(m), "r"(rank));
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(ra), "r"(v), "r"(rm)
               : "memory");
}

// one direction of a (bi)directional layer; both directions run in the same launch (grid.y)
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

// shared memory: h_{t-1} as bf16 hi / lo [2 buffers][kNB][H+8] (row padding: the 8 batch rows of a B fragment
// fall into distinct banks), gates [kNB][128]
template <int H>
struct SmemF {
  uint16_t h_hi[2][kNB * (H + 8)];
  uint16_t h_lo[2][kNB * (H + 8)];
  float g[kNB * kRows];
  uint64_t bar[2];                 // bar[k]: all kNB*H (hi, lo) pairs of h buffer k have arrived
};

// xproj [B,L,4H] (x W_ih^T + b_ih + b_hh), w_hh [4H,H], lengths [B].
// out (pre-zeroed), acts [B,L,4H] (activated i,f,g,o), cs [B,L,H] (cell state), h_last/c_last.
template <int H>
__global__ void __launch_bounds__(kThreads, 1)
lstm_seq_fwd_kernel(DirF d0, DirF d1, const int32_t* __restrict__ lengths, int B, int L, int ld_out, int ld_last, int dbg) {
  constexpr int C = H / kHS;       // cluster size
  constexpr int HP = H + 8;        // padded bf16 row of the h buffers
  constexpr int KS = H / 16;       // k-steps of 16
  const DirF d = blockIdx.y == 0 ? d0 : d1;
  const float* __restrict__ xproj = d.xproj;
  const float* __restrict__ w_hh = d.w_hh;
  float* __restrict__ out = d.out;
  float* __restrict__ acts = d.acts;
  float* __restrict__ cs = d.cs;
  const int reverse = d.reverse;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  SmemF<H>& sm = *reinterpret_cast<SmemF<H>*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rank = (int)cluster_ctarank();
  const int group = blockIdx.x / C;
  const int b0 = group * kNB;

  LSTAMP0(6);
  // resident weight fragments: local row lr = gate*32 + unit  <->  global row gate*H + rank*32 + unit;
  // warp w owns local rows 16w..16w+15 (A operand, row-major 16x16 tiles: a0/a2 row r0, a1/a3 row r0+8)
  const int r0 = lane >> 2, c0 = (lane & 3) * 2;
  uint32_t a_hi[KS][4], a_lo[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int lr = 16 * warp + r0 + (q & 1) * 8, col = ks * 16 + c0 + (q >> 1) * 8;
      const int grow = (lr >> 5) * H + rank * kHS + (lr & 31);
      const float2 wv = __ldg(reinterpret_cast<const float2*>(w_hh + (size_t)grow * H + col));
      split_pair(wv.x, wv.y, a_hi[ks][q], a_lo[ks][q]);
    }
  }
  for (int i = tid; i < 2 * kNB * HP; i += kThreads) {
    (&sm.h_hi[0][0])[i] = 0;
    (&sm.h_lo[0][0])[i] = 0;
  }
  if (tid == 0) {
    mbar_init(&sm.bar[0], 1);
    mbar_init(&sm.bar[1], 1);
    fence_mbar_init();
  }

  // pointwise role: thread = (batch pb, unit pu); c lives in a register for the whole sequence
  const int pb = tid >> 5, pu = tid & 31;
  const int b = b0 + pb;
  const bool valid = b < B;
  const int len = valid ? min(lengths[b], L) : 0;
  int gmax = 0;
  for (int i = 0; i < kNB; ++i)
    if (b0 + i < B) gmax = max(gmax, min(lengths[b0 + i], L));
  float c_reg = 0.f, h_reg = 0.f;
  const int ug = rank * kHS + pu;   // global hidden unit

  cluster_sync_all();               // zeroed h + initialised barriers visible cluster-wide before any DSMEM traffic

  float xp[4] = {0.f, 0.f, 0.f, 0.f};
  auto load_x = [&](int t) {
    if (valid && t >= 0 && t < len) {
      const float* p = xproj + ((size_t)b * L + t) * (4 * H) + ug;
#pragma unroll
      for (int k = 0; k < 4; ++k) xp[k] = __ldg(p + k * H);
    }
  };
  int t = reverse ? gmax - 1 : 0;
  const int dt = reverse ? -1 : 1;
  load_x(t);
  LSTAMP0(7);
  for (int s = 0; s < gmax; ++s, t += dt) {
    const int cur = s & 1, nxt = cur ^ 1;
    LSTAMP(0);
    if (tid == 0) mbar_expect_tx(&sm.bar[nxt], kNB * H * 4);       // this step's h_t: kNB*H (hi, lo) pairs from the C CTAs
    // ---- gates[n][16w + r] = sum_k W[row][k] h[n][k] on the tensor cores (B operand: n = batch row) ----
    {
      // six independent accumulator chains (3 products x even/odd k-step): a single chain would serialise
      // 3*KS dependent MMAs (~25 cycles each) on the step's critical path
      float ac[6][4];
#pragma unroll
      for (int i = 0; i < 6; ++i) ac[i][0] = ac[i][1] = ac[i][2] = ac[i][3] = 0.f;
      const uint16_t* hh = sm.h_hi[cur] + r0 * HP + c0;            // B fragment: n = lane/4, k pair = (lane%4)*2
      const uint16_t* hl = sm.h_lo[cur] + r0 * HP + c0;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(hh + ks * 16), bh1 = *reinterpret_cast<const uint32_t*>(hh + ks * 16 + 8);
        const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(hl + ks * 16), bl1 = *reinterpret_cast<const uint32_t*>(hl + ks * 16 + 8);
        mma_bf16(ac[(ks & 1) * 3 + 0], a_hi[ks], bh0, bh1);
        mma_bf16(ac[(ks & 1) * 3 + 1], a_hi[ks], bl0, bl1);
        mma_bf16(ac[(ks & 1) * 3 + 2], a_lo[ks], bh0, bh1);
      }
      float acc[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = ((ac[0][j] + ac[3][j]) + (ac[1][j] + ac[4][j])) + (ac[2][j] + ac[5][j]);
      // D fragment: acc[0..1] = row r0, batch c0 / c0+1; acc[2..3] = row r0+8
      sm.g[c0 * kRows + 16 * warp + r0] = acc[0];
      sm.g[(c0 + 1) * kRows + 16 * warp + r0] = acc[1];
      sm.g[c0 * kRows + 16 * warp + r0 + 8] = acc[2];
      sm.g[(c0 + 1) * kRows + 16 * warp + r0 + 8] = acc[3];
    }
    LSTAMP(1);
    __syncthreads();
    LSTAMP(2);
    // ---- pointwise for (pb, pu) ----
    const float x0 = xp[0], x1 = xp[1], x2 = xp[2], x3 = xp[3];
    load_x(t + dt);                                  // prefetch next step's input projection
    const bool live = valid && t < len;
    float h_new = h_reg;
    if (live) {
      const float* gp = sm.g + pb * kRows + pu;
      const float ig = sigmoidf_(gp[0] + x0), fg = sigmoidf_(gp[32] + x1), gg = tanhf(gp[64] + x2),
                  og = sigmoidf_(gp[96] + x3);
      c_reg = fg * c_reg + ig * gg;
      h_new = og * tanhf(c_reg);
      h_reg = h_new;
      const size_t o = (size_t)b * L + t;
      out[o * ld_out + ug] = h_new;
      cs[o * H + ug] = c_reg;
      float* ap = acts + o * (4 * H) + ug;
      ap[0] = ig; ap[H] = fg; ap[2 * H] = gg; ap[3 * H] = og;
    }
    LSTAMP(3);
    // ---- all-gather h_t: my unit's value into every CTA's next buffer ----
    // (as bf16 hi/lo: even units send the packed hi pair (pu, pu+1), odd units the packed lo pair (pu-1, pu))
    {
      const uint32_t hi16 = bf16_bits(h_new), lo16 = bf16_bits(h_new - bf16_val(hi16));
      const uint32_t o_hi = __shfl_xor_sync(0xffffffffu, hi16, 1), o_lo = __shfl_xor_sync(0xffffffffu, lo16, 1);
      const bool even = (pu & 1) == 0;
      const uint32_t word = even ? (hi16 | (o_hi << 16)) : (o_lo | (lo16 << 16));
      uint16_t* dst = even ? sm.h_hi[nxt] + pb * HP + ug : sm.h_lo[nxt] + pb * HP + ug - 1;
#pragma unroll
      for (int r = 0; r < C; ++r) dsmem_st_async_u32(dst, &sm.bar[nxt], (uint32_t)r, word);
    }
    // Every CTA (this one included) has delivered its slice once the byte count is reached.  Buffer reuse is
    // safe without a further barrier: a peer writes h[cur] again only in step s+1, which it enters after it
    // received THIS CTA's slice of step s — sent after the matvec above finished reading h[cur].
    LSTAMP(4);
    mbar_wait_cluster(&sm.bar[nxt], (uint32_t)(s >> 1) & 1u);
    LSTAMP(5);
  }
  LSTAMP0(8);
  if (valid) {
    d.h_last[(size_t)b * ld_last + ug] = h_reg;
    d.c_last[(size_t)b * ld_last + ug] = c_reg;
  }
  cluster_sync_all();               // no CTA exits while a peer could still address its shared memory
  LSTAMP0(9);
}

// Backward through time.  d_out [B,L,H] (grad of `out`, may be NULL), d_hlast/d_clast [B,H] (may be NULL).
// d_xproj [B,L,4H] must be pre-zeroed (masked steps stay zero).
template <int H>
struct SmemB {
  uint16_t dg_hi[kNB * (kRows + 8)];     // this CTA's dgates for the group as bf16 hi / lo (B operand, padded rows)
  uint16_t dg_lo[kNB * (kRows + 8)];
  float recv[2][(H / kHS) * kNB * kHS];  // partial dh for my units from every CTA (double buffered over steps)
  uint64_t bar[2];                       // bar[k]: all C*kNB*32 partials of recv[k] have arrived
};

template <int H>
__global__ void __launch_bounds__(kThreads, 1)
lstm_seq_bwd_kernel(DirB d0, DirB d1, const int32_t* __restrict__ lengths, int B, int L, int ld_out, int ld_last, int dbg) {
  constexpr int C = H / kHS;
  constexpr int RP = kRows + 8;    // padded bf16 row of the dgates buffers
  constexpr int MT = H / 128;      // 16-column m-tiles of dh per warp (8 warps cover H columns)
  constexpr int KS = kRows / 16;   // k-steps over this CTA's 128 gate rows
  const DirB d = blockIdx.y == 0 ? d0 : d1;
  const float* __restrict__ w_hh = d.w_hh;
  const float* __restrict__ acts = d.acts;
  const float* __restrict__ cs = d.cs;
  const float* __restrict__ d_out = d.d_out;
  const float* __restrict__ d_hlast = d.d_hlast;
  const float* __restrict__ d_clast = d.d_clast;
  float* __restrict__ d_xproj = d.d_xproj;
  const int reverse = d.reverse;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  SmemB<H>& sm = *reinterpret_cast<SmemB<H>*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rank = (int)cluster_ctarank();
  const int group = blockIdx.x / C;
  const int b0 = group * kNB;
  // resident fragments of W^T restricted to this CTA's gate rows: dh[k][n] = sum_lr W[lr][k] dg[n][lr];
  // A[m = column k][kk = local row lr]; warp w owns columns 16*(MT*w + mt) .. +15
  const int r0 = lane >> 2, c0 = (lane & 3) * 2;
  uint32_t a_hi[MT][KS][4], a_lo[MT][KS][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int k = 16 * (MT * warp + mt) + r0 + (q & 1) * 8, lr = ks * 16 + c0 + (q >> 1) * 8;
        const int g0 = (lr >> 5) * H + rank * kHS + (lr & 31), g1 = ((lr + 1) >> 5) * H + rank * kHS + ((lr + 1) & 31);
        split_pair(__ldg(w_hh + (size_t)g0 * H + k), __ldg(w_hh + (size_t)g1 * H + k), a_hi[mt][ks][q], a_lo[mt][ks][q]);
      }
    }
  }
  const int pb = tid >> 5, pu = tid & 31;
  const int b = b0 + pb;
  const bool valid = b < B;
  const int len = valid ? min(lengths[b], L) : 0;
  int gmax = 0;
  for (int i = 0; i < kNB; ++i)
    if (b0 + i < B) gmax = max(gmax, min(lengths[b0 + i], L));
  const int ug = rank * kHS + pu;
  float dh = (valid && d_hlast) ? d_hlast[(size_t)b * ld_last + ug] : 0.f;   // carried gradient wrt h_t, c_t of my unit
  float dc = (valid && d_clast) ? d_clast[(size_t)b * ld_last + ug] : 0.f;
  if (tid == 0) {
    mbar_init(&sm.bar[0], 1);
    mbar_init(&sm.bar[1], 1);
    fence_mbar_init();
  }
  cluster_sync_all();

  // forward visited t = 0..gmax-1 (or gmax-1..0 when reversed); walk it backwards
  int t = reverse ? 0 : gmax - 1;
  const int dt = reverse ? 1 : -1;
  for (int s = 0; s < gmax; ++s, t += dt) {
    const bool live = valid && t < len;
    if (tid == 0) mbar_expect_tx(&sm.bar[s & 1], C * kNB * kHS * 4);
    float dgi = 0.f, dgf = 0.f, dgg = 0.f, dgo = 0.f;
    if (live) {
      const size_t o = (size_t)b * L + t;
      const float* ap = acts + o * (4 * H) + ug;
      const float ig = ap[0], fg = ap[H], gg = ap[2 * H], og = ap[3 * H];
      const float c1 = cs[o * H + ug];
      const int tp = t - (reverse ? -1 : 1);                        // the step that produced c_{prev}
      const float c0 = (tp >= 0 && tp < len) ? cs[((size_t)b * L + tp) * H + ug] : 0.f;
      const float dht = dh + (d_out ? d_out[o * ld_out + ug] : 0.f);
      const float tc = tanhf(c1);
      const float dct = dc + dht * og * (1.f - tc * tc);
      dgi = dct * gg * ig * (1.f - ig);
      dgf = dct * c0 * fg * (1.f - fg);
      dgg = dct * ig * (1.f - gg * gg);
      dgo = dht * tc * og * (1.f - og);
      dc = dct * fg;
      float* dp = d_xproj + o * (4 * H) + ug;
      dp[0] = dgi; dp[H] = dgf; dp[2 * H] = dgg; dp[3 * H] = dgo;
    }
    {
      const float dgv[4] = {dgi, dgf, dgg, dgo};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t hi = bf16_bits(dgv[k]);
        sm.dg_hi[pb * RP + k * kHS + pu] = (uint16_t)hi;
        sm.dg_lo[pb * RP + k * kHS + pu] = (uint16_t)bf16_bits(dgv[k] - bf16_val(hi));
      }
    }
    __syncthreads();
    // partial dh_prev[n][k] over my 128 rows on the tensor cores (B operand: n = batch row)
    {
      float ac[MT][3][4];                                      // independent chains per product (x MT tiles)
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int i = 0; i < 3; ++i) ac[mt][i][0] = ac[mt][i][1] = ac[mt][i][2] = ac[mt][i][3] = 0.f;
      const uint16_t* gh = sm.dg_hi + r0 * RP + c0;
      const uint16_t* gl = sm.dg_lo + r0 * RP + c0;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(gh + ks * 16), bh1 = *reinterpret_cast<const uint32_t*>(gh + ks * 16 + 8);
        const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(gl + ks * 16), bl1 = *reinterpret_cast<const uint32_t*>(gl + ks * 16 + 8);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          mma_bf16(ac[mt][0], a_hi[mt][ks], bh0, bh1);
          mma_bf16(ac[mt][1], a_hi[mt][ks], bl0, bl1);
          mma_bf16(ac[mt][2], a_lo[mt][ks], bh0, bh1);
        }
      }
      float acc[MT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[mt][j] = (ac[mt][0][j] + ac[mt][1][j]) + ac[mt][2][j];
      // reduce-scatter: column k belongs to CTA k/32, slot [my rank][n][k%32]
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int k = 16 * (MT * warp + mt) + r0 + (q >> 1) * 8, n = c0 + (q & 1);
          dsmem_st_async_f32(sm.recv[s & 1] + (rank * kNB + n) * kHS + (k & 31), &sm.bar[s & 1], (uint32_t)(k >> 5),
                             acc[mt][q]);
        }
      }
    }
    mbar_wait_cluster(&sm.bar[s & 1], (uint32_t)(s >> 1) & 1u);
    // owner: dh_{t-1}[pb][my unit] = sum over the C partials (live rows); frozen rows pass dh through
    if (live) {
      float v = 0.f;
#pragma unroll
      for (int r = 0; r < C; ++r) v += sm.recv[s & 1][(r * kNB + pb) * kHS + pu];
      dh = v;
    }
    // recv is double buffered: a peer's stores of step s+1 go to the other half, and its stores of step
    // s+2 come after it received this CTA's partials of step s+1, which are sent only after the reads above.
    // sm.dg_* are rewritten in step s+1 only after this wait, i.e. after every local thread finished its MMAs.
  }
  cluster_sync_all();               // no CTA exits while a peer could still address its shared memory
}

template <typename K, typename D>
int launch_cluster(K kernel, size_t smem, int C, int n_dir, int B, cudaStream_t stream, D d0, D d1, const int32_t* lengths,
                   int L, int ld_out, int ld_last) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(((B + kNB - 1) / kNB) * C, n_dir);
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

template <int H>
int launch_fwd(DirF d0, DirF d1, int n_dir, const int32_t* lengths, int B, int L, int ld_out, int ld_last,
               cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    VLN_CHECK_CUDA(cudaFuncSetAttribute(lstm_seq_fwd_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)sizeof(SmemF<H>)));
    configured = true;
  }
  return launch_cluster(lstm_seq_fwd_kernel<H>, sizeof(SmemF<H>), H / kHS, n_dir, B, stream, d0, d1, lengths, L, ld_out,
                        ld_last);
}

template <int H>
int launch_bwd(DirB d0, DirB d1, int n_dir, const int32_t* lengths, int B, int L, int ld_out, int ld_last,
               cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    VLN_CHECK_CUDA(cudaFuncSetAttribute(lstm_seq_bwd_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)sizeof(SmemB<H>)));
    configured = true;
  }
  return launch_cluster(lstm_seq_bwd_kernel<H>, sizeof(SmemB<H>), H / kHS, n_dir, B, stream, d0, d1, lengths, L, ld_out,
                        ld_last);
}

}  // namespace

// How many clusters of the recurrence kernels can be resident at once (cudaOccupancyMaxActiveClusters): a launch of
// (B/8) * n_dir clusters beyond this runs in two waves, i.e. twice the 80-step latency chain.
template <typename K>
static int max_clusters(K kernel, size_t smem, int C, int* out) {
  VLN_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
extern "C" int vln_debug_lstm_occupancy(int H, int* fwd_clusters, int* bwd_clusters) {
  VLN_REQUIRE(fwd_clusters && bwd_clusters, "bad arguments");
  if (H == 256) {
    if (int rc = max_clusters(lstm_seq_fwd_kernel<256>, sizeof(SmemF<256>), 8, fwd_clusters)) return rc;
    return max_clusters(lstm_seq_bwd_kernel<256>, sizeof(SmemB<256>), 8, bwd_clusters);
  }
  if (H == 128) {
    if (int rc = max_clusters(lstm_seq_fwd_kernel<128>, sizeof(SmemF<128>), 4, fwd_clusters)) return rc;
    return max_clusters(lstm_seq_bwd_kernel<128>, sizeof(SmemB<128>), 4, bwd_clusters);
  }
  vln_set_error("vln_debug_lstm_occupancy: hidden size %d per direction is not supported (128 or 256)", H);
  return -1;
}

extern "C" int vln_debug_lstm_stamps(unsigned long long* out_host /*[8]*/) {
  VLN_CHECK_CUDA(cudaDeviceSynchronize());
  VLN_CHECK_CUDA(cudaMemcpyFromSymbol(out_host, g_lstm_stamps, sizeof(unsigned long long) * 12));
  return 0;
}

// tcgen05 version of the two kernels (csrc/lstm_tc.cu): the default; VLN_LSTM_VARIANT=mma keeps the mma.sync kernels above.
int vln_lstm_tc_fwd(const float* const* xproj, const float* const* w_hh, const int32_t* lengths, float* out,
                    float* const* acts, float* const* cs, float* h_last, float* c_last, int B, int L, int H, int n_dir,
                    cudaStream_t stream);
int vln_lstm_tc_bwd(const float* const* w_hh, const int32_t* lengths, const float* const* acts, const float* const* cs,
                    const float* d_out, const float* d_hlast, const float* d_clast, float* const* d_xproj, int B, int L,
                    int H, int n_dir, cudaStream_t stream);
static int g_variant = -1;                                // -1: read VLN_LSTM_VARIANT once; 0: mma.sync; 1: tcgen05
static bool use_tc() {
  if (g_variant < 0) {
    const char* e = getenv("VLN_LSTM_VARIANT");
    g_variant = (e && !strcmp(e, "mma")) ? 0 : 1;
  }
  return g_variant == 1;
}
extern "C" int vln_lstm_set_variant(int variant) {
  VLN_REQUIRE(variant >= -1 && variant <= 1, "variant must be -1 (environment), 0 (mma.sync) or 1 (tcgen05)");
  g_variant = variant;
  return 0;
}

// n_dir = 1 or 2.  Direction k uses xproj[k], w_hh[k], acts[k], cs[k] and writes columns [k*H, (k+1)*H) of
// out [B,L,n_dir*H] / h_last / c_last [B,n_dir*H]; direction 1 runs reversed in time.
extern "C" int vln_lstm_seq_fwd(const float* const* xproj, const float* const* w_hh, const int32_t* lengths, float* out,
                                float* const* acts, float* const* cs, float* h_last, float* c_last, int B, int L, int H,
                                int n_dir, void* stream) {
  VLN_REQUIRE(xproj && w_hh && lengths && out && acts && cs && h_last && c_last && B > 0 && L > 0, "bad arguments");
  VLN_REQUIRE(n_dir == 1 || n_dir == 2, "n_dir must be 1 or 2");
  DirF d[2] = {};
  for (int k = 0; k < n_dir; ++k) {
    VLN_REQUIRE(xproj[k] && w_hh[k] && acts[k] && cs[k], "null per-direction pointer");
    VLN_REQUIRE(((uintptr_t)w_hh[k] & 15) == 0, "w_hh must be 16-byte aligned");
    d[k] = DirF{xproj[k], w_hh[k], out + k * H, acts[k], cs[k], h_last + k * H, c_last + k * H, k};
  }
  if ((use_tc() && (H == 256 || H == 128)) || H == 512)        // 512 per direction exists on the tcgen05 path only
    return vln_lstm_tc_fwd(xproj, w_hh, lengths, out, acts, cs, h_last, c_last, B, L, H, n_dir, (cudaStream_t)stream);
  if (H == 256) return launch_fwd<256>(d[0], d[1], n_dir, lengths, B, L, n_dir * H, n_dir * H, (cudaStream_t)stream);
  if (H == 128) return launch_fwd<128>(d[0], d[1], n_dir, lengths, B, L, n_dir * H, n_dir * H, (cudaStream_t)stream);
  vln_set_error("vln_lstm_seq_fwd: hidden size %d per direction is not supported (128 or 256)", H);
  return -1;
}

// d_out [B,L,n_dir*H] (nullable), d_hlast / d_clast [B,n_dir*H] (nullable); d_xproj[k] [B,L,4H] pre-zeroed.
extern "C" int vln_lstm_seq_bwd(const float* const* w_hh, const int32_t* lengths, const float* const* acts,
                                const float* const* cs, const float* d_out, const float* d_hlast, const float* d_clast,
                                float* const* d_xproj, int B, int L, int H, int n_dir, void* stream) {
  VLN_REQUIRE(w_hh && lengths && acts && cs && d_xproj && B > 0 && L > 0, "bad arguments");
  VLN_REQUIRE(n_dir == 1 || n_dir == 2, "n_dir must be 1 or 2");
  DirB d[2] = {};
  for (int k = 0; k < n_dir; ++k) {
    VLN_REQUIRE(w_hh[k] && acts[k] && cs[k] && d_xproj[k], "null per-direction pointer");
    d[k] = DirB{w_hh[k], acts[k], cs[k], d_out ? d_out + k * H : nullptr, d_hlast ? d_hlast + k * H : nullptr,
                d_clast ? d_clast + k * H : nullptr, d_xproj[k], k};
  }
  if ((use_tc() && (H == 256 || H == 128)) || H == 512)
    return vln_lstm_tc_bwd(w_hh, lengths, acts, cs, d_out, d_hlast, d_clast, d_xproj, B, L, H, n_dir, (cudaStream_t)stream);
  if (H == 256) return launch_bwd<256>(d[0], d[1], n_dir, lengths, B, L, n_dir * H, n_dir * H, (cudaStream_t)stream);
  if (H == 128) return launch_bwd<128>(d[0], d[1], n_dir, lengths, B, L, n_dir * H, n_dir * H, (cudaStream_t)stream);
  vln_set_error("vln_lstm_seq_bwd: hidden size %d per direction is not supported (128 or 256)", H);
  return -1;
}
