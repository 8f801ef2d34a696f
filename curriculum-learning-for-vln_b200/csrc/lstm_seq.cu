// Persistent length-masked LSTM recurrence (one direction of EncoderLSTM's nn.LSTM, units.py:58-71),
// forward and backward-through-time, one launch each.
//
// The reference packs the padded batch and calls cuDNN; here a THREAD-BLOCK CLUSTER owns a group
// of kNB batch rows for the whole sequence:
//   * CTA r of the cluster owns hidden units [32r, 32r+32): its 128 gate rows (i,f,g,o x 32) of
//     W_hh stay resident in shared memory for all timesteps (128 x H fp32, padded rows);
//   * every step each CTA computes its 128 gate pre-activations for the group's kNB rows from the
//     full h_{t-1} (kept in shared memory), applies the LSTM pointwise update to its 32 units
//     (c stays in registers), and broadcasts its slice of h_t into every peer's shared memory
//     through DSMEM; one cluster barrier per step;
//   * rows stop at their own length (reverse rows start there) exactly like a packed sequence:
//     state frozen and output zero where t >= length;  steps past the group's longest row are skipped.
// Backward runs the same structure in reverse time: the pointwise gradient for the CTA's units,
// partial dh_{t-1} = dgates . W_hh over its 128 rows, reduce-scattered to the owning CTAs via DSMEM.
// dgates (= d xproj) goes to global memory; dW_hh / dW_ih / dx are single large GEMMs outside.
//
// Roofline class: latency (80 serial steps); FMA work per step per CTA = kNB*128*H.
#include "common.cuh"

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
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
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

// shared memory: W slice [128][H+8], h double buffer [2][kNB][H], gates [kNB][128]
template <int H>
struct SmemF {
  float w[kRows * (H + 8)];
  float h[2][kNB * H];
  float g[kNB * kRows];
  uint64_t bar[2];                 // bar[k]: all kNB*H values of h buffer k have arrived
};

// xproj [B,L,4H] (x W_ih^T + b_ih + b_hh), w_hh [4H,H], lengths [B].
// out (pre-zeroed), acts [B,L,4H] (activated i,f,g,o), cs [B,L,H] (cell state), h_last/c_last.
template <int H>
__global__ void __launch_bounds__(kThreads, 1)
lstm_seq_fwd_kernel(DirF d0, DirF d1, const int32_t* __restrict__ lengths, int B, int L, int ld_out, int ld_last) {
  constexpr int C = H / kHS;       // cluster size
  constexpr int WS = H + 8;        // padded row stride (floats): rows 0..3 x k-phase 0/1 hit distinct 16 B slots
  const DirF d = blockIdx.y == 0 ? d0 : d1;
  const float* __restrict__ xproj = d.xproj;
  const float* __restrict__ w_hh = d.w_hh;
  float* __restrict__ out = d.out;
  float* __restrict__ acts = d.acts;
  float* __restrict__ cs = d.cs;
  const int reverse = d.reverse;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  SmemF<H>& sm = *reinterpret_cast<SmemF<H>*>(smem_raw);
  const int tid = threadIdx.x;
  const int rank = (int)cluster_ctarank();
  const int group = blockIdx.x / C;
  const int b0 = group * kNB;

  // resident weight slice: local row lr = gate*32 + unit  <->  global row gate*H + rank*32 + unit
  for (int i = tid; i < kRows * (H / 4); i += kThreads) {
    const int lr = i / (H / 4), c4 = i - lr * (H / 4);
    const int grow = (lr >> 5) * H + rank * kHS + (lr & 31);
    reinterpret_cast<float4*>(sm.w + lr * WS)[c4] = __ldg(reinterpret_cast<const float4*>(w_hh + (size_t)grow * H) + c4);
  }
  for (int i = tid; i < 2 * kNB * H; i += kThreads) (&sm.h[0][0])[i] = 0.f;
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
  // matvec role: thread = (row mr, k-half mk)
  const int mr = tid >> 1, mk = tid & 1;                  // the two k-phases interleave in 16-byte steps
  const float* wrow = sm.w + mr * WS + mk * 4;

  cluster_sync_all();               // weights + zeroed h visible cluster-wide before any DSMEM traffic

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
  for (int s = 0; s < gmax; ++s, t += dt) {
    const int cur = s & 1, nxt = cur ^ 1;
    if (tid == 0) mbar_expect_tx(&sm.bar[nxt], kNB * H * 4);       // this step's h_t: kNB*H floats from the C CTAs
    // ---- gates[b][row] partial sums over this thread's half of K ----
    float acc[kNB];
#pragma unroll
    for (int i = 0; i < kNB; ++i) acc[i] = 0.f;
    const float* hb = sm.h[cur] + mk * 4;
#pragma unroll 4
    for (int k = 0; k < H; k += 8) {
      const float4 w4 = *reinterpret_cast<const float4*>(wrow + k);
#pragma unroll
      for (int i = 0; i < kNB; ++i) {
        const float4 h4 = *reinterpret_cast<const float4*>(hb + i * H + k);
        acc[i] = fmaf(w4.x, h4.x, fmaf(w4.y, h4.y, fmaf(w4.z, h4.z, fmaf(w4.w, h4.w, acc[i]))));
      }
    }
#pragma unroll
    for (int i = 0; i < kNB; ++i) {
      acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 1);
      if (mk == 0) sm.g[i * kRows + mr] = acc[i];
    }
    __syncthreads();
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
    // ---- all-gather h_t: my unit's value into every CTA's next buffer ----
    float* dst = sm.h[nxt] + pb * H + ug;
#pragma unroll
    for (int r = 0; r < C; ++r) dsmem_st_async_f32(dst, &sm.bar[nxt], (uint32_t)r, h_new);
    // Every CTA (this one included) has delivered its slice once the byte count is reached.  Buffer reuse is
    // safe without a further barrier: a peer writes h[cur] again only in step s+1, which it enters after it
    // received THIS CTA's slice of step s — sent after the matvec above finished reading h[cur].
    mbar_wait_cluster(&sm.bar[nxt], (uint32_t)(s >> 1) & 1u);
  }
  if (valid) {
    d.h_last[(size_t)b * ld_last + ug] = h_reg;
    d.c_last[(size_t)b * ld_last + ug] = c_reg;
  }
  cluster_sync_all();               // no CTA exits while a peer could still address its shared memory
}

// Backward through time.  d_out [B,L,H] (grad of `out`, may be NULL), d_hlast/d_clast [B,H] (may be NULL).
// d_xproj [B,L,4H] must be pre-zeroed (masked steps stay zero).
template <int H>
struct SmemB {
  float w[kRows * (H + 8)];
  float dg[kNB * kRows];                 // this CTA's dgates for the group
  float recv[2][(H / kHS) * kNB * kHS];  // partial dh for my units from every CTA (double buffered over steps)
  uint64_t bar[2];                       // bar[k]: all C*kNB*32 partials of recv[k] have arrived
};

template <int H>
__global__ void __launch_bounds__(kThreads, 1)
lstm_seq_bwd_kernel(DirB d0, DirB d1, const int32_t* __restrict__ lengths, int B, int L, int ld_out, int ld_last) {
  constexpr int C = H / kHS;
  constexpr int WS = H + 8;
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
  const int tid = threadIdx.x;
  const int rank = (int)cluster_ctarank();
  const int group = blockIdx.x / C;
  const int b0 = group * kNB;
  for (int i = tid; i < kRows * (H / 4); i += kThreads) {
    const int lr = i / (H / 4), c4 = i - lr * (H / 4);
    const int grow = (lr >> 5) * H + rank * kHS + (lr & 31);
    reinterpret_cast<float4*>(sm.w + lr * WS)[c4] = __ldg(reinterpret_cast<const float4*>(w_hh + (size_t)grow * H) + c4);
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
    float* gp = sm.dg + pb * kRows + pu;
    gp[0] = dgi; gp[32] = dgf; gp[64] = dgg; gp[96] = dgo;
    __syncthreads();
    // partial dh_prev[bb][k] over my 128 rows: thread tid owns column k = tid (H <= 256) for all kNB rows
    if (tid < H) {
      float acc[kNB];
#pragma unroll
      for (int i = 0; i < kNB; ++i) acc[i] = 0.f;
#pragma unroll 4
      for (int lr = 0; lr < kRows; ++lr) {
        const float w = sm.w[lr * WS + tid];
#pragma unroll
        for (int i = 0; i < kNB; ++i) acc[i] = fmaf(sm.dg[i * kRows + lr], w, acc[i]);
      }
      // reduce-scatter: column k belongs to CTA k/32, slot [my rank][bb][k%32]
      const uint32_t owner = (uint32_t)(tid >> 5);
#pragma unroll
      for (int i = 0; i < kNB; ++i)
        dsmem_st_async_f32(sm.recv[s & 1] + (rank * kNB + i) * kHS + (tid & 31), &sm.bar[s & 1], owner, acc[i]);
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
    // sm.dg is rewritten in step s+1 only after this wait, i.e. after every local thread finished its matvec.
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
  VLN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, d0, d1, lengths, B, L, ld_out, ld_last));
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
    d[k] = DirF{xproj[k], w_hh[k], out + k * H, acts[k], cs[k], h_last + k * H, c_last + k * H, k};
  }
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
  if (H == 256) return launch_bwd<256>(d[0], d[1], n_dir, lengths, B, L, n_dir * H, n_dir * H, (cudaStream_t)stream);
  if (H == 128) return launch_bwd<128>(d[0], d[1], n_dir, lengths, B, L, n_dir * H, n_dir * H, (cudaStream_t)stream);
  vln_set_error("vln_lstm_seq_bwd: hidden size %d per direction is not supported (128 or 256)", H);
  return -1;
}
