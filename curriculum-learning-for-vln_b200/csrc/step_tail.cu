// Tail of a fused EnvDrop decoder step in ONE launch: candidate logits + action head + simulator transition + the next
// pass's action embedding (policy.py:199-206 with envdrop.py:166-219 and policy.py:222-223).
//
// Same arithmetic, in the same order, as cand_logits_fwd_kernel (pano.cu) followed by policy_env_act_kernel (step.cu)
// — the results are bit-identical (tests/test_kernels_gpu.py) — but one launch instead of two, and the chain of
// dependent index loads of the transition is flattened:
//   * one CTA per episode, 16 warps: warp j scores candidate slot j straight from the table rows (L2 hits: the panorama
//     attention of this step just read them);
//   * while the rows are in flight, warp 0 already holds everything the transition needs that does not depend on the
//     action: the whole candidate row of the current viewpoint (one lane per slot), goal, ended flag, distance, target;
//   * after the action is known the new viewpoint's tables are fetched by all lanes at once, and the teacher slot of
//     the NEXT step is one ballot over the lanes instead of a serial search with a dependent load per candidate.
#include "common.cuh"

namespace {

__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }
__device__ __forceinline__ uint4 apply_keep(uint4 x, const Philox8& r, uint32_t thr) {
  uint32_t* w = reinterpret_cast<uint32_t*>(&x);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t m = (philox_keep(r, 2 * i, thr) ? 0x0000FFFFu : 0u) | (philox_keep(r, 2 * i + 1, thr) ? 0xFFFF0000u : 0u);
    w[i] &= m;
  }
  return x;
}
// drop(tanh(W_a angle128(view) + b_a)) with the four group sums of every weight row (step.cu: act_embed_one)
__device__ __forceinline__ float act_embed_one(const float* __restrict__ wg_row, const float* __restrict__ p4, float bias) {
  const float4 w = __ldg(reinterpret_cast<const float4*>(wg_row)), p = __ldg(reinterpret_cast<const float4*>(p4));
  return tanhf(fmaf(p.w, w.w, fmaf(p.z, w.z, fmaf(p.y, w.y, fmaf(p.x, w.x, bias)))));
}

// (measured: no gain — 4.45 vs 4.44 ms per iteration; the gather of the next step is not what the chain waits for)
constexpr bool kPrefetchNext = false;

struct TailArgs {
  // candidate logits
  const __nv_bfloat16* table; const int32_t* vp; const int32_t* view; const float* cand_ang4; const float* tgt;
  float* logits; float drop_p; uint64_t off_cand;
  // action head
  const int32_t* target; int feedback; const uint64_t* rng; uint64_t off_sample;
  float* ce; int32_t* action; float* logp; float* entropy; float* probs;
  // simulator transition
  const uint8_t* ended_in; const float* dist_in; const int32_t* goal;
  const int32_t* cand_vp; const int32_t* cand_view; const int32_t* n_cand; const int32_t* next_hop;
  const float* dist_tbl; const int64_t* sq_off; const int32_t* vp_local;
  int32_t* vp_out; int32_t* view_out; uint8_t* ended_out; float* dist_out; int32_t* teacher_out;
  float* reward; float* mask; int32_t* n_active;
  // next pass's action embedding
  const float* pose4; const float* w_act; const float* b_act; float* act; float* xh; int ld_xh; int E; float p_act;
  uint64_t off_act;
};

__global__ void __launch_bounds__(512) cand_policy_env_act_kernel(const __grid_constant__ TailArgs a, int B,
                                                                  const __grid_constant__ ChainLink link) {
  __shared__ __align__(16) float ts[VLN_F];
  __shared__ float ta[4];
  __shared__ float s_logit[VLN_NSLOT];
  __shared__ __align__(16) float s_pose[VLN_V * 4];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, j = tid >> 5;
  CHAIN_BEGIN(a.rng, 6);
  pdl_trigger();
  // Everything up to the wait reads data that was complete long before the preceding kernel (the GEMM that produces
  // tgt) started: this step's state was written by the previous step's tail, five grid-wide kernels back (each of
  // which had to drain before the next could become resident), and the tables are constant.  So the index chain, the
  // candidate rows and the whole speculative transition are in flight while that GEMM is still finishing.
  const int g = a.vp[b];
  const int vw_in = a.view[b];
  const int n = a.n_cand[g];
  // ---- warp 0: the transition, SPECULATIVELY for every action it could take.  Lane j < n stands for "move to
  //      candidate j", every lane >= n for "stay" (STOP / ended / ignored): each lane fetches the tables of ITS next
  //      viewpoint — distance to the goal, next hop, the teacher slot there — before the dependency wait, so that once
  //      the action is known the outcome is a handful of shuffles instead of two dependent global round trips ----
  int row_view = 0, sp_cur = g, sp_n2 = 0, sp_teach = 0, gl = 0, tg = -1;
  float sp_d = 0.f, d_in = 0.f;
  bool was_ended = false;
  if (j == 0) {
    if (lane < n) {
      row_view = a.cand_view[(size_t)g * VLN_CMAX + lane];
      sp_cur = a.cand_vp[(size_t)g * VLN_CMAX + lane];
    }
    gl = a.goal[b];
    tg = a.target ? a.target[b] : -1;
    was_ended = a.ended_in[b] != 0;
    d_in = a.dist_in[b];
    const int vloc_g = a.vp_local[gl];
    const int64_t so = a.sq_off[sp_cur];
    sp_n2 = a.n_cand[sp_cur];
    int c2[VLN_CMAX];
#pragma unroll
    for (int k = 0; k < VLN_CMAX; ++k) c2[k] = a.cand_vp[(size_t)sp_cur * VLN_CMAX + k];
    sp_d = a.dist_tbl[so + vloc_g];
    const int nh = a.next_hop[so + vloc_g];
    sp_teach = sp_n2;                                          // base.py:174-177: STOP when no candidate leads on
#pragma unroll
    for (int k = VLN_CMAX - 1; k >= 0; --k)
      if (k < sp_n2 && c2[k] == nh) sp_teach = k;              // first matching slot
    if (sp_cur == gl) sp_teach = sp_n2;
  }
  // the rest of what the post-action part reads and that no predecessor writes: the pose table, the sampling uniform,
  // this lane's two rows of the action-embedding weights and their dropout keep-scales
  if (tid >= 32 && tid < 32 + VLN_V && a.xh != nullptr)
    reinterpret_cast<float4*>(s_pose)[tid - 32] = __ldg(reinterpret_cast<const float4*>(a.pose4) + (tid - 32));
  float u_sample = 0.f;
  float4 wa[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
  float ba[2] = {0.f, 0.f}, ka[2] = {1.f, 1.f};
  if (j == 0) {
    if ((a.feedback & 3) == 2) u_sample = philox_uniform(philox8(a.rng[0], a.rng[1] + a.off_sample, (uint64_t)b), 0);
    if (a.xh != nullptr) {
      const bool on = a.p_act > 0.f;
      const uint32_t thr_a = drop_threshold(a.p_act);
      const float sc_a = on ? 1.0f / (1.0f - a.p_act) : 1.0f;
      const uint64_t seed_a = on ? a.rng[0] : 0, off_a = on ? a.rng[1] + a.off_act : 0;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int q = lane + 32 * r;
        if (q < a.E) {
          wa[r] = __ldg(reinterpret_cast<const float4*>(a.w_act + (size_t)q * 4));
          ba[r] = a.b_act[q];
          const int i = b * a.E + q;
          if (on) ka[r] = philox_keep(philox8(seed_a, off_a, (uint64_t)(i >> 3)), i & 7, thr_a) ? sc_a : 0.f;
        }
      }
    }
  }
  // this warp's candidate row (8 x 16 bytes per lane), its angle feature, and the feature-dropout keep mask
  // (policy.py:226-231) applied in registers — none of it depends on the preceding kernel
  uint4 xr[8];
  float4 an = make_float4(0.f, 0.f, 0.f, 0.f);
  if (j < n) {
    const int cv = a.cand_view[(size_t)g * VLN_CMAX + j];
    an = __ldg(reinterpret_cast<const float4*>(a.cand_ang4 + (((size_t)g * VLN_CMAX + j) * 12 + (vw_in % 12)) * 4));
    const uint4* src = reinterpret_cast<const uint4*>(a.table + ((size_t)g * VLN_V + cv) * VLN_IMG);
#pragma unroll
    for (int it = 0; it < 8; ++it) xr[it] = __ldg(src + it * 32 + lane);
    if (a.drop_p > 0.f) {
      const uint32_t thr = drop_threshold(a.drop_p);
      const uint64_t seed = a.rng[0], offset = a.rng[1] + a.off_cand;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const uint64_t e = (((uint64_t)b * VLN_NSLOT + j) * VLN_IMG + (uint64_t)(it * 32 + lane) * 8) >> 3;
        xr[it] = apply_keep(xr[it], philox8(seed, offset, e), thr);
      }
    }
  }
  chain_wait_cta(link);
  CHAIN_MARK(2);
  // ---- candidate logits (cand_logits_fwd_kernel) ----
  for (int i = tid; i < VLN_F; i += 512) ts[i] = a.tgt[(size_t)b * VLN_F + i];
  __syncthreads();
  if (j < 4) {
    float s = warp_sum(ts[VLN_IMG + 32 * j + lane]);
    if (lane == 0) ta[j] = s;
  }
  __syncthreads();
  float res;
  if (j < n) {
    float acc = 0.f;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int vi = it * 32 + lane;
      const uint4 x = xr[it];
      const float4 q0 = reinterpret_cast<const float4*>(ts)[vi * 2], q1 = reinterpret_cast<const float4*>(ts)[vi * 2 + 1];
      acc += bf16lo(x.x) * q0.x + bf16hi(x.x) * q0.y + bf16lo(x.y) * q0.z + bf16hi(x.y) * q0.w +
             bf16lo(x.z) * q1.x + bf16hi(x.z) * q1.y + bf16lo(x.w) * q1.z + bf16hi(x.w) * q1.w;
    }
    acc = warp_sum(acc);
    if (a.drop_p > 0.f) acc *= 1.0f / (1.0f - a.drop_p);
    res = acc + an.x * ta[0] + an.y * ta[1] + an.z * ta[2] + an.w * ta[3] + 0.f;
  } else if (j == n) {
    res = 0.f;                                   // END slot: all-zero feature row (base.py:152-153)
  } else {
    res = -INFINITY;                             // length2mask + masked_fill_(-inf)
  }
  if (lane == 0) {
    a.logits[(size_t)b * VLN_NSLOT + j] = res;
    s_logit[j] = res;
  }
  __syncthreads();
  if (j != 0) {                                                // (one arrival per warp: 16 per episode)
    chain_signal_warp(link);
    return;
  }

  // ---- action head (policy_fwd_kernel / policy_env_act_kernel) ----
  const float x = lane < VLN_NSLOT ? s_logit[lane] : -INFINITY;
  const float m = warp_max(x);
  const float e = (x == -INFINITY) ? 0.f : expf(x - m);
  const float s = warp_sum(e);
  const float p = e / s;
  const float lp = x - m - logf(s);
  const float ent = -warp_sum(p > 0.f ? p * lp : 0.f);
  const int t_from = a.feedback >> 8;
  const int mode = (t_from > 0 && b >= t_from - 1) ? 0 : (a.feedback & 3);
  int act_id;
  if (mode == 0) {
    act_id = tg;
  } else if (mode == 1) {
    act_id = __ffs(__ballot_sync(0xffffffffu, x == m)) - 1;
  } else {
    const float u = u_sample;
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
  if (lane < VLN_NSLOT) a.probs[(size_t)b * VLN_NSLOT + lane] = p;

  // ---- simulator transition (env_step_kernel): pick the speculated outcome of the chosen action ----
  const bool stop = was_ended || act_id < 0 || act_id >= n;
  const int sel = stop ? 31 : act_id;                          // lane 31 always speculated "stay" (n <= 15)
  const int cur = __shfl_sync(0xffffffffu, sp_cur, sel);
  const int nv = __shfl_sync(0xffffffffu, row_view, sel);
  const int vw = stop ? vw_in : nv;
  const float d = __shfl_sync(0xffffffffu, sp_d, sel);
  const int teach = __shfl_sync(0xffffffffu, sp_teach, sel);
  // The next step's panorama attention gathers the 36 rows of the new viewpoint (147 456 contiguous bytes) as soon as it
  // has waited for this grid: ask L2 for them now, so that HBM round trip runs under this kernel's tail, the grid
  // hand-over and the attention kernel's index load instead of after them.
  if (a.xh != nullptr && !stop && kPrefetchNext) {
    const char* nxt = reinterpret_cast<const char*>(a.table + (size_t)cur * VLN_V * VLN_IMG) + (size_t)lane * (VLN_V * VLN_IMG * 2 / 32);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nxt), "r"(VLN_V * VLN_IMG * 2 / 32) : "memory");
  }
  if (lane == 0) {
    a.ce[b] = tg >= 0 ? -lp_t : 0.f;
    a.action[b] = act_id;
    a.logp[b] = act_id >= 0 ? lp_a : 0.f;
    a.entropy[b] = ent;
    a.vp_out[b] = cur;
    a.view_out[b] = vw;
    float r = 0.f;
    if (!was_ended) {
      if (stop) r = d < 3.0f ? 2.f : -2.f;
      else { const float dd = d_in - d; r = dd > 0.f ? 1.f : (dd < 0.f ? -1.f : 0.f); }
    }
    a.reward[b] = r;
    a.mask[b] = was_ended ? 0.f : 1.f;
    a.dist_out[b] = d;
    const bool now_ended = was_ended || stop;
    a.ended_out[b] = now_ended ? 1 : 0;
    a.teacher_out[b] = now_ended ? -1 : teach;
    if (a.n_active && !now_ended) atomicAdd(a.n_active, 1);
  }
  // ---- next pass's action embedding for the new view (act_fwd_kernel; weights, bias, keep-scales fetched above) ----
  if (a.xh == nullptr) {
    chain_signal_warp(link);
    CHAIN_MARK(3);
    return;
  }
  const float4 pz = *reinterpret_cast<const float4*>(s_pose + vw * 4);
  if (a.E <= 64) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int q = lane + 32 * r;
      if (q < a.E) {
        const float v = tanhf(fmaf(pz.w, wa[r].w, fmaf(pz.z, wa[r].z, fmaf(pz.y, wa[r].y, fmaf(pz.x, wa[r].x, ba[r])))));
        a.act[b * a.E + q] = v;
        a.xh[(size_t)b * a.ld_xh + q] = v * ka[r];
      }
    }
  } else {                                       // wider embeddings: the general loop
    const bool on = a.p_act > 0.f;
    const uint32_t thr_a = drop_threshold(a.p_act);
    const float sc_a = on ? 1.0f / (1.0f - a.p_act) : 1.0f;
    const uint64_t seed_a = on ? a.rng[0] : 0, off_a = on ? a.rng[1] + a.off_act : 0;
    for (int q = lane; q < a.E; q += 32) {
      const float v = act_embed_one(a.w_act + (size_t)q * 4, a.pose4 + (size_t)vw * 4, a.b_act[q]);
      const int i = b * a.E + q;
      a.act[i] = v;
      float k = 1.f;
      if (on) k = philox_keep(philox8(seed_a, off_a, (uint64_t)(i >> 3)), i & 7, thr_a) ? sc_a : 0.f;
      a.xh[(size_t)b * a.ld_xh + q] = v * k;
    }
  }
  chain_signal_warp(link);
  CHAIN_MARK(3);
}

}  // namespace

extern "C" int vln_cand_policy_env_act_fwd(
    const vln_ctx* ctx, const int32_t* vp_in, const int32_t* view_in, const float* cand_ang4, const float* tgt, float* logits,
    float drop_p, uint64_t off_cand, const int32_t* target, int feedback, const uint64_t* rng, uint64_t off_sample, float* ce,
    int32_t* action, float* logp, float* entropy, float* probs, const uint8_t* ended_in, const float* dist_in,
    const int32_t* goal, const int32_t* cand_vp, const int32_t* cand_view, const int32_t* n_cand, const int32_t* next_hop,
    const float* dist_tbl, const int64_t* sq_off, const int32_t* vp_local, int32_t* vp_out, int32_t* view_out,
    uint8_t* ended_out, float* dist_out, int32_t* teacher_out, float* reward, float* mask, int32_t* n_active,
    const float* pose4, const float* w_act, const float* b_act, float* act, float* xh, int ld_xh, int E, float p_act,
    uint64_t off_act, int B, void* stream) {
  VLN_REQUIRE(ctx && vp_in && view_in && cand_ang4 && tgt && logits && ce && action && logp && entropy && probs && ended_in &&
                  dist_in && goal && cand_vp && cand_view && n_cand && next_hop && dist_tbl && sq_off && vp_local && vp_out &&
                  view_out && ended_out && dist_out && teacher_out && reward && mask && B > 0,
              "bad arguments");
  VLN_REQUIRE(feedback >= 0 && (feedback & 3) <= 2 && ((feedback & 3) != 0 || target) && ((feedback & 3) != 2 || rng) &&
                  ((feedback >> 8) == 0 || target),
              "bad feedback mode");
  VLN_REQUIRE(drop_p >= 0.f && drop_p < 1.f && (drop_p == 0.f || rng), "feature dropout needs 0 <= p < 1 and an rng state");
  VLN_REQUIRE(!xh || (pose4 && w_act && b_act && act && E > 0 && (p_act == 0.f || rng)), "action embedding needs its weights");
  TailArgs a;
  a.table = ctx->table; a.vp = vp_in; a.view = view_in; a.cand_ang4 = cand_ang4; a.tgt = tgt; a.logits = logits;
  a.drop_p = drop_p; a.off_cand = off_cand;
  a.target = target; a.feedback = feedback; a.rng = rng; a.off_sample = off_sample;
  a.ce = ce; a.action = action; a.logp = logp; a.entropy = entropy; a.probs = probs;
  a.ended_in = ended_in; a.dist_in = dist_in; a.goal = goal;
  a.cand_vp = cand_vp; a.cand_view = cand_view; a.n_cand = n_cand; a.next_hop = next_hop; a.dist_tbl = dist_tbl;
  a.sq_off = sq_off; a.vp_local = vp_local;
  a.vp_out = vp_out; a.view_out = view_out; a.ended_out = ended_out; a.dist_out = dist_out; a.teacher_out = teacher_out;
  a.reward = reward; a.mask = mask; a.n_active = n_active;
  a.pose4 = pose4; a.w_act = w_act; a.b_act = b_act; a.act = act; a.xh = xh; a.ld_xh = ld_xh; a.E = E; a.p_act = p_act;
  a.off_act = off_act;
  const ChainLink link = vln_chain_link((cudaStream_t)stream, (unsigned int)B * 16u);
  VLN_CHECK_CUDA(vln_launch_linked(cand_policy_env_act_kernel, dim3(B), dim3(512), 0, (cudaStream_t)stream, a, B, link));
  return 0;
}
