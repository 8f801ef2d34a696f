// Glue kernels of the fused EnvDrop decoder step (EnvDropDecoder.forward policy.py:208-246 inside
// the rollout loop envdrop.py:134-221) and of its backward.
//
// The step is a chain of grid-wide dependencies (skinny GEMM -> gather/attention -> GEMM -> ...);
// everything BETWEEN two such kernels — tanh, the four nn.Dropout sites, the concatenations
// torch.cat((prev_act_emb, visual_feat)) / cat((weighted, h)), the action embedding, the LSTMCell
// pointwise half, the action head + simulator transition — is folded into the handful of kernels
// below, writing straight into the row-strided operand buffers of the next GEMM:
//
//   XH[t]  [B, 64 + 2176 + 512]  = [ drop(tanh(W_a pose + b_a)) | visual attention output | h~_{t-1} ]
//   WH[t]  [B, 512 + 512]        = [ text-attention weighted context | drop(h_t) ]
//
// Dropout masks come from the same (seed, base + call_off, element/8) Philox blocks as
// vln_dropout on a dense [B, n] tensor, so vln_dropout_mask reproduces them for the oracle.
// All kernels here are latency-bound (a few thousand threads); the point is the launch count.
#include "common.cuh"

namespace {

struct Drop {
  uint64_t seed, offset;
  uint32_t thr;
  float scale;
  bool on;
};

__device__ __forceinline__ Drop make_drop(float p, const uint64_t* rng, uint64_t call_off) {
  Drop d;
  d.on = p > 0.f;
  d.thr = drop_threshold(p);
  d.scale = d.on ? 1.0f / (1.0f - p) : 1.0f;
  d.seed = d.on ? rng[0] : 0;
  d.offset = d.on ? rng[1] + call_off : 0;
  return d;
}
// keep-scale factors (0 or 1/(1-p)) of the 8 elements of Philox block `blk`
__device__ __forceinline__ void keep8(const Drop& d, uint64_t blk, float (&k)[8]) {
  if (!d.on) {
#pragma unroll
    for (int j = 0; j < 8; ++j) k[j] = 1.f;
    return;
  }
  const Philox8 r = philox8(d.seed, d.offset, blk);
#pragma unroll
  for (int j = 0; j < 8; ++j) k[j] = philox_keep(r, j, d.thr) ? d.scale : 0.f;
}
__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// The four per-element kernels between the step's GEMMs (state_fwd/bwd, lstm_pw_drop_fwd/bwd) sit on the step's
// latency chain with only B*H = 32 K elements: a thread takes kE = 2 consecutive elements (four threads share one
// 8-element Philox block and each recomputes it), so the dependent chain of transcendentals per thread is a quarter
// of the 8-per-thread layout the wider kernels use (4.5 -> ~2.5 us for the LSTM pointwise kernel at B = 64).
constexpr int kE = 2;
__device__ __forceinline__ void keepE(const Drop& d, uint64_t blk, int sub, float (&k)[kE]) {
  if (!d.on) {
#pragma unroll
    for (int j = 0; j < kE; ++j) k[j] = 1.f;
    return;
  }
  const Philox8 r = philox8(d.seed, d.offset, blk);
#pragma unroll
  for (int j = 0; j < kE; ++j) k[j] = philox_keep(r, sub + j, d.thr) ? d.scale : 0.f;
}
__device__ __forceinline__ void ldE(const float* p, float (&v)[kE]) {
  const float2 a = *reinterpret_cast<const float2*>(p);
  v[0] = a.x; v[1] = a.y;
}
__device__ __forceinline__ void stE(float* p, const float (&v)[kE]) { *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]); }
// thread -> (row b, first column h, Philox block of the dense [B,H] tensor, lane offset inside the block)
#define ELEM_INDEX()                                                   \
  const int i_ = blockIdx.x * blockDim.x + threadIdx.x;               \
  const int per_row = H / kE;                                          \
  if (i_ >= B * per_row) return;                                       \
  const int b = i_ / per_row, h = (i_ - b * per_row) * kE;             \
  const uint64_t i = ((uint64_t)b * H + h) >> 3;                       \
  const int sub = h & 7

// ---- h~ state: tanh + the two dropout sites that consume it -------------------------------------
__device__ __forceinline__ void state_fwd_body(const float* __restrict__ src, int apply_tanh, float* __restrict__ xh_next,
                                               int ld_xh, float* __restrict__ hq_next, float* __restrict__ hc_cur, int B, int H,
                                               float p, const uint64_t* __restrict__ rng, uint64_t off_q, uint64_t off_c) {
  ELEM_INDEX();
  float v[kE], k[kE], o[kE];
  ldE(src + (size_t)b * H + h, v);
  if (apply_tanh) {
#pragma unroll
    for (int j = 0; j < kE; ++j) v[j] = tanhf(v[j]);
  }
  if (xh_next) stE(xh_next + (size_t)b * ld_xh + h, v);
  if (hq_next) {
    keepE(make_drop(p, rng, off_q), i, sub, k);
#pragma unroll
    for (int j = 0; j < kE; ++j) o[j] = v[j] * k[j];
    stE(hq_next + (size_t)b * H + h, o);
  }
  if (hc_cur) {
    keepE(make_drop(p, rng, off_c), i, sub, k);
#pragma unroll
    for (int j = 0; j < kE; ++j) o[j] = v[j] * k[j];
    stE(hc_cur + (size_t)b * H + h, o);
  }
}

__global__ void state_fwd_kernel(const float* __restrict__ src, int apply_tanh, float* __restrict__ xh_next, int ld_xh,
                                 float* __restrict__ hq_next, float* __restrict__ hc_cur, int B, int H, float p,
                                 const uint64_t* __restrict__ rng, uint64_t off_q, uint64_t off_c,
                                 const __grid_constant__ ChainLink link) {
  pdl_trigger();
  chain_wait_cta(link);
  state_fwd_body(src, apply_tanh, xh_next, ld_xh, hq_next, hc_cur, B, H, p, rng, off_q, off_c);
  chain_signal_cta(link);
}

// d_src = (drop_c'(d_hc) + d_xh_next + drop_q'(d_hq_next)) * (apply_tanh ? 1 - h~^2 : 1)
__device__ __forceinline__ void state_bwd_body(const float* __restrict__ d_hc, const float* __restrict__ d_xh_next, int ld_dxh,
                                               const float* __restrict__ d_hq_next, const float* __restrict__ htilde, int ld_h,
                                               int apply_tanh, float* __restrict__ d_src, int B, int H, float p,
                                               const uint64_t* __restrict__ rng, uint64_t off_q, uint64_t off_c) {
  ELEM_INDEX();
  float g[kE] = {0.f, 0.f}, v[kE], k[kE];
  if (d_hc) {
    ldE(d_hc + (size_t)b * H + h, v);
    keepE(make_drop(p, rng, off_c), i, sub, k);
#pragma unroll
    for (int j = 0; j < kE; ++j) g[j] += v[j] * k[j];
  }
  if (d_xh_next) {
    ldE(d_xh_next + (size_t)b * ld_dxh + h, v);
#pragma unroll
    for (int j = 0; j < kE; ++j) g[j] += v[j];
  }
  if (d_hq_next) {
    ldE(d_hq_next + (size_t)b * H + h, v);
    keepE(make_drop(p, rng, off_q), i, sub, k);
#pragma unroll
    for (int j = 0; j < kE; ++j) g[j] += v[j] * k[j];
  }
  if (apply_tanh) {
    ldE(htilde + (size_t)b * ld_h + h, v);
#pragma unroll
    for (int j = 0; j < kE; ++j) g[j] *= 1.f - v[j] * v[j];
  }
  stE(d_src + (size_t)b * H + h, g);
}
__global__ void state_bwd_kernel(const float* __restrict__ d_hc, const float* __restrict__ d_xh_next, int ld_dxh,
                                 const float* __restrict__ d_hq_next, const float* __restrict__ htilde, int ld_h,
                                 int apply_tanh, float* __restrict__ d_src, int B, int H, float p,
                                 const uint64_t* __restrict__ rng, uint64_t off_q, uint64_t off_c,
                                 const __grid_constant__ ChainLink link) {
  pdl_trigger();
  chain_wait_cta(link);
  state_bwd_body(d_hc, d_xh_next, ld_dxh, d_hq_next, htilde, ld_h, apply_tanh, d_src, B, H, p, rng, off_q, off_c);
  chain_signal_cta(link);
}

// ---- action embedding of the agent's pose: drop(tanh(W_a angle128(view) + b_a)) ------------------
// angle128 = pose4[view] each value repeated x32 (misc.py:286-293), so the 128-wide dot product is
// 4 values against 4 group sums of the weight row.
// `wg` [E,4] holds the four group sums of every weight row (wg[j][k] = sum_i W[j][32k+i], computed once
// per rollout on the host side of the C-ABI)
__device__ __forceinline__ float act_embed_one(const float* __restrict__ wg_row, const float* __restrict__ p4, float bias) {
  const float4 w = __ldg(reinterpret_cast<const float4*>(wg_row)), p = __ldg(reinterpret_cast<const float4*>(p4));
  return tanhf(fmaf(p.w, w.w, fmaf(p.z, w.z, fmaf(p.y, w.y, fmaf(p.x, w.x, bias)))));
}

__global__ void act_fwd_kernel(const int32_t* __restrict__ view, const float* __restrict__ pose4,
                               const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ act,
                               float* __restrict__ xh, int ld_xh, int B, int E, float p,
                               const uint64_t* __restrict__ rng, uint64_t call_off) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * E) return;
  const int b = i / E, j = i - b * E;
  const float a = act_embed_one(w + (size_t)j * 4, pose4 + (size_t)view[b] * 4, bias[j]);
  act[i] = a;
  float k[8];
  keep8(make_drop(p, rng, call_off), (uint64_t)(i >> 3), k);
  xh[(size_t)b * ld_xh + j] = a * k[i & 7];
}

// d_actpre[t,b,j] = drop'(d_xh[t,b,j]) * (1 - act^2) for n_steps steps at once (step t draws its mask from
// stream call_off + t*off_stride; d_xh rows of step t start at d_xh + t*B*ld_dxh)
__global__ void act_bwd_kernel(const float* __restrict__ d_xh, int ld_dxh, const float* __restrict__ act,
                               float* __restrict__ d_actpre, int B, int E, int n_steps, float p,
                               const uint64_t* __restrict__ rng, uint64_t call_off, uint64_t off_stride) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_steps * B * E) return;
  const int t = g / (B * E), i = g - t * (B * E);
  const int b = i / E, j = i - b * E;
  float k[8];
  keep8(make_drop(p, rng, call_off + (uint64_t)t * off_stride), (uint64_t)(i >> 3), k);
  const float a = act[g];
  d_actpre[g] = d_xh[((size_t)t * B + b) * ld_dxh + j] * k[i & 7] * (1.f - a * a);
}

// ---- action head + simulator transition + next step's action embedding: one warp per episode ------
// (vln_policy_fwd, vln_env_step and vln_envdrop_act_fwd back to back; same arithmetic, one launch)
struct EnvTables {
  const int32_t* cand_vp; const int32_t* cand_view; const int32_t* n_cand; const int32_t* next_hop;
  const float* dist_tbl; const int64_t* sq_off; const int32_t* vp_local;
};
__device__ __forceinline__ int teacher_slot_(int cur, int goal, const EnvTables& e) {
  const int n = e.n_cand[cur];
  if (cur == goal) return n;
  const int nh = e.next_hop[e.sq_off[cur] + e.vp_local[goal]];
  for (int j = 0; j < n; ++j)
    if (e.cand_vp[(size_t)cur * VLN_CMAX + j] == nh) return j;
  return n;
}

__global__ void policy_env_act_kernel(const float* __restrict__ logits, const int32_t* __restrict__ target, int feedback,
                                      const uint64_t* __restrict__ rng, uint64_t off_sample, float* __restrict__ ce,
                                      int32_t* __restrict__ action, float* __restrict__ logp, float* __restrict__ entropy,
                                      float* __restrict__ probs, const int32_t* __restrict__ vp_in,
                                      const int32_t* __restrict__ view_in, const uint8_t* __restrict__ ended_in,
                                      const float* __restrict__ dist_in, const int32_t* __restrict__ goal, EnvTables env,
                                      int32_t* __restrict__ vp_out, int32_t* __restrict__ view_out,
                                      uint8_t* __restrict__ ended_out, float* __restrict__ dist_out,
                                      int32_t* __restrict__ teacher_out, float* __restrict__ reward,
                                      float* __restrict__ mask, int32_t* __restrict__ n_active,
                                      const float* __restrict__ pose4, const float* __restrict__ w_act,
                                      const float* __restrict__ b_act, float* __restrict__ act, float* __restrict__ xh,
                                      int ld_xh, int E, float p_act, uint64_t off_act, int B) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  // ---- action head (policy_fwd_kernel) ----
  const float x = lane < VLN_NSLOT ? logits[(size_t)b * VLN_NSLOT + lane] : -INFINITY;
  const float m = warp_max(x);
  const float e = (x == -INFINITY) ? 0.f : expf(x - m);
  const float s = warp_sum(e);
  const float p = e / s;
  const float lp = x - m - logf(s);
  const float ent = -warp_sum(p > 0.f ? p * lp : 0.f);
  const int tg = target ? target[b] : -1;
  // feedback = mode | (teacher_from + 1) << 8: rows >= teacher_from follow the teacher whatever the mode (the
  // teacher-forced and the sampled rollout of one EnvDrop iteration stepped as one launch, trainer.py:411-421)
  const int t_from = feedback >> 8;
  const int mode = (t_from > 0 && b >= t_from - 1) ? 0 : (feedback & 3);
  int act_id;
  if (mode == 0) {
    act_id = tg;
  } else if (mode == 1) {
    act_id = __ffs(__ballot_sync(0xffffffffu, x == m)) - 1;
  } else {
    const float u = philox_uniform(philox8(rng[0], rng[1] + off_sample, (uint64_t)b), 0);
    float cdf = p;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(0xffffffffu, cdf, o);
      if (lane >= o) cdf += t;
    }
    const unsigned hit = __ballot_sync(0xffffffffu, p > 0.f && cdf > u);
    const unsigned valid = __ballot_sync(0xffffffffu, p > 0.f);
    act_id = hit ? __ffs(hit) - 1 : 31 - __clz(valid);
  }
  const float lp_t = __shfl_sync(0xffffffffu, lp, tg >= 0 ? tg : 0);
  const float lp_a = __shfl_sync(0xffffffffu, lp, act_id >= 0 ? act_id : 0);
  if (lane < VLN_NSLOT) probs[(size_t)b * VLN_NSLOT + lane] = p;
  // ---- simulator transition (env_step_kernel), lane 0 ----
  int vw = 0;
  if (lane == 0) {
    ce[b] = tg >= 0 ? -lp_t : 0.f;
    action[b] = act_id;
    logp[b] = act_id >= 0 ? lp_a : 0.f;
    entropy[b] = ent;
    int cur = vp_in[b];
    vw = view_in[b];
    const int g = goal[b];
    const bool was_ended = ended_in[b] != 0;
    const bool stop = was_ended || act_id < 0 || act_id >= env.n_cand[cur];
    if (!stop) {
      vw = env.cand_view[(size_t)cur * VLN_CMAX + act_id];
      cur = env.cand_vp[(size_t)cur * VLN_CMAX + act_id];
    }
    vp_out[b] = cur;
    view_out[b] = vw;
    const float d = env.dist_tbl[env.sq_off[cur] + env.vp_local[g]];
    float r = 0.f;
    if (!was_ended) {
      if (stop) r = d < 3.0f ? 2.f : -2.f;
      else { const float dd = dist_in[b] - d; r = dd > 0.f ? 1.f : (dd < 0.f ? -1.f : 0.f); }
    }
    reward[b] = r;
    mask[b] = was_ended ? 0.f : 1.f;
    dist_out[b] = d;
    const bool now_ended = was_ended || stop;
    ended_out[b] = now_ended ? 1 : 0;
    teacher_out[b] = now_ended ? -1 : teacher_slot_(cur, g, env);
    if (n_active && !now_ended) atomicAdd(n_active, 1);
  }
  // ---- next step's action embedding (act_fwd_kernel) for the new view ----
  if (xh == nullptr) return;
  vw = __shfl_sync(0xffffffffu, vw, 0);
  const Drop dr = make_drop(p_act, rng, off_act);
  for (int j = lane; j < E; j += 32) {
    const float a = act_embed_one(w_act + (size_t)j * 4, pose4 + (size_t)vw * 4, b_act[j]);
    const int i = b * E + j;
    act[i] = a;
    float k[8];
    keep8(dr, (uint64_t)(i >> 3), k);
    xh[(size_t)b * ld_xh + j] = a * k[i & 7];
  }
}

// ---- nn.LSTMCell pointwise half + dropout of h_1 into the text-attention operand buffer -----------
__global__ void lstm_pw_drop_fwd_kernel(const float* __restrict__ gates, const float* __restrict__ c0,
                                        float* __restrict__ h1, float* __restrict__ c1, float* __restrict__ acts,
                                        float* __restrict__ h1_drop, int ld_drop, int B, int H, float p,
                                        const uint64_t* __restrict__ rng, uint64_t call_off) {
  pdl_trigger();
  pdl_wait();
  ELEM_INDEX();
  const float* gr = gates + (size_t)b * 4 * H + h;
  float gi[kE], gf[kE], gg[kE], go[kE], c[kE], hh[kE], k[kE];
  ldE(gr, gi); ldE(gr + H, gf); ldE(gr + 2 * H, gg); ldE(gr + 3 * H, go);
  ldE(c0 + (size_t)b * H + h, c);
#pragma unroll
  for (int j = 0; j < kE; ++j) {
    gi[j] = sigmoidf_(gi[j]); gf[j] = sigmoidf_(gf[j]); gg[j] = tanhf(gg[j]); go[j] = sigmoidf_(go[j]);
    c[j] = gf[j] * c[j] + gi[j] * gg[j];
    hh[j] = go[j] * tanhf(c[j]);
  }
  stE(c1 + (size_t)b * H + h, c);
  stE(h1 + (size_t)b * H + h, hh);
  if (acts) {
    float* ar = acts + (size_t)b * 4 * H + h;
    stE(ar, gi); stE(ar + H, gf); stE(ar + 2 * H, gg); stE(ar + 3 * H, go);
  }
  if (h1_drop) {
    keepE(make_drop(p, rng, call_off), i, sub, k);
#pragma unroll
    for (int j = 0; j < kE; ++j) hh[j] *= k[j];
    stE(h1_drop + (size_t)b * ld_drop + h, hh);
  }
}

// d_h1 = drop'(d_h1_drop) + d_h1_extra ; then the LSTMCell pointwise backward
__global__ void lstm_pw_drop_bwd_kernel(const float* __restrict__ acts, const float* __restrict__ c0,
                                        const float* __restrict__ c1, const float* __restrict__ d_h1_drop, int ld_drop,
                                        const float* __restrict__ d_h1_extra, const float* __restrict__ d_c1,
                                        float* __restrict__ d_gates, float* __restrict__ d_c0, int B, int H, float p,
                                        const uint64_t* __restrict__ rng, uint64_t call_off) {
  pdl_trigger();
  pdl_wait();
  ELEM_INDEX();
  const float* ar = acts + (size_t)b * 4 * H + h;
  float gi[kE], gf[kE], gg[kE], go[kE], cp[kE], cn[kE], dh[kE], dc[kE], k[kE], v[kE];
  ldE(ar, gi); ldE(ar + H, gf); ldE(ar + 2 * H, gg); ldE(ar + 3 * H, go);
  ldE(c0 + (size_t)b * H + h, cp);
  ldE(c1 + (size_t)b * H + h, cn);
#pragma unroll
  for (int j = 0; j < kE; ++j) dh[j] = 0.f;
  if (d_h1_drop) {
    ldE(d_h1_drop + (size_t)b * ld_drop + h, v);
    keepE(make_drop(p, rng, call_off), i, sub, k);
#pragma unroll
    for (int j = 0; j < kE; ++j) dh[j] = v[j] * k[j];
  }
  if (d_h1_extra) {
    ldE(d_h1_extra + (size_t)b * H + h, v);
#pragma unroll
    for (int j = 0; j < kE; ++j) dh[j] += v[j];
  }
  if (d_c1) ldE(d_c1 + (size_t)b * H + h, dc);
  else {
#pragma unroll
    for (int j = 0; j < kE; ++j) dc[j] = 0.f;
  }
  float di[kE], df[kE], dg[kE], dgo[kE];
#pragma unroll
  for (int j = 0; j < kE; ++j) {
    const float tc = tanhf(cn[j]);
    const float dcj = dc[j] + dh[j] * go[j] * (1.f - tc * tc);
    di[j] = dcj * gg[j] * gi[j] * (1.f - gi[j]);
    df[j] = dcj * cp[j] * gf[j] * (1.f - gf[j]);
    dg[j] = dcj * gi[j] * (1.f - gg[j] * gg[j]);
    dgo[j] = dh[j] * tc * go[j] * (1.f - go[j]);
    dc[j] = dcj * gf[j];
  }
  float* dr = d_gates + (size_t)b * 4 * H + h;
  stE(dr, di); stE(dr + H, df); stE(dr + 2 * H, dg); stE(dr + 3 * H, dgo);
  stE(d_c0 + (size_t)b * H + h, dc);
}

}  // namespace

#define STREAM ((cudaStream_t)stream)

extern "C" int vln_envdrop_state_fwd(const float* src, int apply_tanh, float* xh_next, int ld_xh, float* hq_next,
                                     float* hc_cur, int B, int H, float p, const uint64_t* rng, uint64_t off_q,
                                     uint64_t off_c, void* stream) {
  VLN_REQUIRE(src && B > 0 && H > 0 && H % 8 == 0, "bad arguments");
  VLN_REQUIRE(p >= 0.f && p < 1.f && (p == 0.f || rng), "dropout needs 0 <= p < 1 and an rng state");
  VLN_REQUIRE(!xh_next || ld_xh % 4 == 0, "xh rows must be 16-byte aligned");
  const int n = B * (H / kE);                              // one thread per kE consecutive elements
  const ChainLink link = vln_chain_link(STREAM, (unsigned int)((n + 127) / 128));
  VLN_CHECK_CUDA(vln_launch_linked(state_fwd_kernel, dim3((n + 127) / 128), dim3(128), 0, STREAM, src, apply_tanh, xh_next,
                                   ld_xh, hq_next, hc_cur, B, H, p, rng, off_q, off_c, link));
  return 0;
}

extern "C" int vln_envdrop_state_bwd(const float* d_hc, const float* d_xh_next, int ld_dxh, const float* d_hq_next,
                                     const float* htilde, int ld_h, int apply_tanh, float* d_src, int B, int H, float p,
                                     const uint64_t* rng, uint64_t off_q, uint64_t off_c, void* stream) {
  VLN_REQUIRE(d_src && B > 0 && H > 0 && H % 8 == 0, "bad arguments");
  VLN_REQUIRE(!apply_tanh || htilde, "tanh backward needs the saved h~");
  VLN_REQUIRE(p >= 0.f && p < 1.f && (p == 0.f || rng), "dropout needs 0 <= p < 1 and an rng state");
  const int n = B * (H / kE);                              // one thread per kE consecutive elements
  const ChainLink link = vln_chain_link(STREAM, (unsigned int)((n + 127) / 128));
  VLN_CHECK_CUDA(vln_launch_linked(state_bwd_kernel, dim3((n + 127) / 128), dim3(128), 0, STREAM, d_hc, d_xh_next, ld_dxh,
                                   d_hq_next, htilde, ld_h, apply_tanh, d_src, B, H, p, rng, off_q, off_c, link));
  return 0;
}

extern "C" int vln_envdrop_act_fwd(const int32_t* view, const float* pose4, const float* w, const float* bias, float* act,
                                   float* xh, int ld_xh, int B, int E, float p, const uint64_t* rng, uint64_t call_off,
                                   void* stream) {
  VLN_REQUIRE(view && pose4 && w && bias && act && xh && B > 0 && E > 0, "bad arguments");
  VLN_REQUIRE(p >= 0.f && p < 1.f && (p == 0.f || rng), "dropout needs 0 <= p < 1 and an rng state");
  act_fwd_kernel<<<(B * E + 127) / 128, 128, 0, STREAM>>>(view, pose4, w, bias, act, xh, ld_xh, B, E, p, rng, call_off);
  VLN_LAUNCH_OK();
  return 0;
}

extern "C" int vln_envdrop_act_bwd(const float* d_xh, int ld_dxh, const float* act, float* d_actpre, int B, int E,
                                   int n_steps, float p, const uint64_t* rng, uint64_t call_off, uint64_t off_stride,
                                   void* stream) {
  VLN_REQUIRE(d_xh && act && d_actpre && B > 0 && E > 0 && n_steps > 0, "bad arguments");
  VLN_REQUIRE(p >= 0.f && p < 1.f && (p == 0.f || rng), "dropout needs 0 <= p < 1 and an rng state");
  act_bwd_kernel<<<(n_steps * B * E + 255) / 256, 256, 0, STREAM>>>(d_xh, ld_dxh, act, d_actpre, B, E, n_steps, p, rng,
                                                                    call_off, off_stride);
  VLN_LAUNCH_OK();
  return 0;
}

extern "C" int vln_policy_env_act_fwd(const float* logits, const int32_t* target, int feedback, const uint64_t* rng,
                                      uint64_t off_sample, float* ce, int32_t* action, float* logp, float* entropy,
                                      float* probs, const int32_t* vp_in, const int32_t* view_in, const uint8_t* ended_in,
                                      const float* dist_in, const int32_t* goal, const int32_t* cand_vp,
                                      const int32_t* cand_view, const int32_t* n_cand, const int32_t* next_hop,
                                      const float* dist_tbl, const int64_t* sq_off, const int32_t* vp_local,
                                      int32_t* vp_out, int32_t* view_out, uint8_t* ended_out, float* dist_out,
                                      int32_t* teacher_out, float* reward, float* mask, int32_t* n_active,
                                      const float* pose4, const float* w_act, const float* b_act, float* act, float* xh,
                                      int ld_xh, int E, float p_act, uint64_t off_act, int B, void* stream) {
  VLN_REQUIRE(logits && ce && action && logp && entropy && probs && vp_in && view_in && ended_in && dist_in && goal &&
                  cand_vp && cand_view && n_cand && next_hop && dist_tbl && sq_off && vp_local && vp_out && view_out &&
                  ended_out && dist_out && teacher_out && reward && mask && B > 0,
              "bad arguments");
  VLN_REQUIRE(feedback >= 0 && (feedback & 3) <= 2 && ((feedback & 3) != 0 || target) && ((feedback & 3) != 2 || rng) &&
                  ((feedback >> 8) == 0 || target),
              "bad feedback mode");
  VLN_REQUIRE(!xh || (pose4 && w_act && b_act && act && E > 0 && (p_act == 0.f || rng)), "action embedding needs its weights");
  EnvTables env{cand_vp, cand_view, n_cand, next_hop, dist_tbl, sq_off, vp_local};
  VLN_CHECK_CUDA(vln_launch_chain(policy_env_act_kernel, dim3((B + 3) / 4), dim3(128), 0, STREAM, logits, target, feedback,
                                  rng, off_sample, ce, action, logp, entropy, probs, vp_in, view_in, ended_in, dist_in,
                                  goal, env, vp_out, view_out, ended_out, dist_out, teacher_out, reward, mask, n_active,
                                  pose4, w_act, b_act, act, xh, ld_xh, E, p_act, off_act, B));
  return 0;
}

extern "C" int vln_lstm_pointwise_drop_fwd(const float* gates, const float* c0, float* h1, float* c1, float* acts,
                                           float* h1_drop, int ld_drop, int B, int H, float p, const uint64_t* rng,
                                           uint64_t call_off, void* stream) {
  VLN_REQUIRE(gates && c0 && h1 && c1 && B > 0 && H > 0 && H % 8 == 0, "bad arguments");
  VLN_REQUIRE(p >= 0.f && p < 1.f && (p == 0.f || rng || !h1_drop), "dropout needs 0 <= p < 1 and an rng state");
  const int n = B * (H / kE);                              // one thread per kE consecutive elements
  VLN_CHECK_CUDA(vln_launch_chain(lstm_pw_drop_fwd_kernel, dim3((n + 127) / 128), dim3(128), 0, STREAM, gates, c0, h1, c1,
                                  acts, h1_drop, ld_drop, B, H, p, rng, call_off));
  return 0;
}

extern "C" int vln_lstm_pointwise_drop_bwd(const float* acts, const float* c0, const float* c1, const float* d_h1_drop,
                                           int ld_drop, const float* d_h1_extra, const float* d_c1, float* d_gates,
                                           float* d_c0, int B, int H, float p, const uint64_t* rng, uint64_t call_off,
                                           void* stream) {
  VLN_REQUIRE(acts && c0 && c1 && d_gates && d_c0 && B > 0 && H > 0 && H % 8 == 0, "bad arguments");
  VLN_REQUIRE(p >= 0.f && p < 1.f && (p == 0.f || rng || !d_h1_drop), "dropout needs 0 <= p < 1 and an rng state");
  const int n = B * (H / kE);                              // one thread per kE consecutive elements
  VLN_CHECK_CUDA(vln_launch_chain(lstm_pw_drop_bwd_kernel, dim3((n + 127) / 128), dim3(128), 0, STREAM, acts, c0, c1,
                                  d_h1_drop, ld_drop, d_h1_extra, d_c1, d_gates, d_c0, B, H, p, rng, call_off));
  return 0;
}
