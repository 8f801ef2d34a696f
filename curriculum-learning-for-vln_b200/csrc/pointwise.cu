// Small fused kernels: LSTM pointwise halves, the action head (masked log-softmax / CE / sampling /
// entropy) with its backward, Philox dropout, the index-table environment step, and the fused
// clip + optimiser pass.  All are launch/latency-bound at B=64; they exist to keep the rollout
// on the device and to collapse ~20 ATen launches per decoder step into one each.
#include "common.cuh"

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---- nn.LSTMCell pointwise (gate order i, f, g, o) ----------------------------------------------
__global__ void lstm_pw_fwd_kernel(const float* __restrict__ gates, const float* __restrict__ c0,
                                   float* __restrict__ h1, float* __restrict__ c1, float* __restrict__ acts, int B,
                                   int H) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H) return;
  const int b = idx / H, h = idx - b * H;
  const float* gr = gates + (size_t)b * 4 * H;
  const float i = sigmoidf_(gr[h]), f = sigmoidf_(gr[H + h]), g = tanhf(gr[2 * H + h]), o = sigmoidf_(gr[3 * H + h]);
  const float c = f * c0[idx] + i * g;
  c1[idx] = c;
  h1[idx] = o * tanhf(c);
  if (acts) {
    float* ar = acts + (size_t)b * 4 * H;
    ar[h] = i; ar[H + h] = f; ar[2 * H + h] = g; ar[3 * H + h] = o;
  }
}

__global__ void lstm_pw_bwd_kernel(const float* __restrict__ acts, const float* __restrict__ c0,
                                   const float* __restrict__ c1, const float* __restrict__ d_h1,
                                   const float* __restrict__ d_c1, float* __restrict__ d_gates,
                                   float* __restrict__ d_c0, int B, int H) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H) return;
  const int b = idx / H, h = idx - b * H;
  const float* ar = acts + (size_t)b * 4 * H;
  const float i = ar[h], f = ar[H + h], g = ar[2 * H + h], o = ar[3 * H + h];
  const float tc = tanhf(c1[idx]);
  const float dh = d_h1 ? d_h1[idx] : 0.f;
  const float dc = (d_c1 ? d_c1[idx] : 0.f) + dh * o * (1.f - tc * tc);
  float* dg = d_gates + (size_t)b * 4 * H;
  dg[h] = dc * g * i * (1.f - i);
  dg[H + h] = dc * c0[idx] * f * (1.f - f);
  dg[2 * H + h] = dc * i * (1.f - g * g);
  dg[3 * H + h] = dh * tc * o * (1.f - o);
  d_c0[idx] = dc * f;
}

// ---- action head: one warp per episode, 16 slots -------------------------------------------------
__global__ void policy_fwd_kernel(const float* __restrict__ logits, const int32_t* __restrict__ target, int feedback,
                                  const uint64_t* __restrict__ rng, uint64_t call_off, float* __restrict__ ce, int32_t* __restrict__ action,
                                  float* __restrict__ logp, float* __restrict__ entropy, float* __restrict__ probs,
                                  int B) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  const float x = lane < VLN_NSLOT ? logits[(size_t)b * VLN_NSLOT + lane] : -INFINITY;
  const float m = warp_max(x);
  const float e = (x == -INFINITY) ? 0.f : expf(x - m);
  const float s = warp_sum(e);
  const float p = e / s;
  const float lp = x - m - logf(s);                       // log-softmax (−inf on masked slots)
  const float ent = -warp_sum(p > 0.f ? p * lp : 0.f);
  const int tg = target ? target[b] : -1;
  int act;
  if (feedback == 0) {
    act = tg;
  } else if (feedback == 1) {                             // first index attaining the maximum
    const unsigned hit = __ballot_sync(0xffffffffu, x == m);
    act = __ffs(hit) - 1;
  } else {                                                // inverse-CDF sample, u ~ Philox(seed, offset, b)
    const float u = philox_uniform(philox8(rng[0], rng[1] + call_off, (uint64_t)b), 0);
    float cdf = p;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(0xffffffffu, cdf, o);
      if (lane >= o) cdf += t;
    }
    const unsigned hit = __ballot_sync(0xffffffffu, p > 0.f && cdf > u);
    const unsigned valid = __ballot_sync(0xffffffffu, p > 0.f);
    act = hit ? __ffs(hit) - 1 : 31 - __clz(valid);
  }
  const float lp_t = __shfl_sync(0xffffffffu, lp, tg >= 0 ? tg : 0);
  const float lp_a = __shfl_sync(0xffffffffu, lp, act >= 0 ? act : 0);
  if (lane < VLN_NSLOT && probs) probs[(size_t)b * VLN_NSLOT + lane] = p;
  if (lane == 0) {
    if (ce) ce[b] = tg >= 0 ? -lp_t : 0.f;              // CrossEntropyLoss(ignore_index=-1, reduction='none')
    if (action) action[b] = act;
    if (logp) logp[b] = act >= 0 ? lp_a : 0.f;
    if (entropy) entropy[b] = ent;
  }
}

__global__ void policy_bwd_kernel(const float* __restrict__ probs, const int32_t* __restrict__ target,
                                  const int32_t* __restrict__ action, const float* __restrict__ entropy,
                                  const float* __restrict__ g_ce, const float* __restrict__ g_logp,
                                  const float* __restrict__ g_ent, float* __restrict__ dlogits, int B) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * VLN_NSLOT) return;
  const int b = idx / VLN_NSLOT, j = idx - b * VLN_NSLOT;
  const float p = probs[idx];
  float d = 0.f;
  if (g_ce && target[b] >= 0) d += g_ce[b] * (p - (j == target[b] ? 1.f : 0.f));
  if (g_logp && action[b] >= 0) d += g_logp[b] * ((j == action[b] ? 1.f : 0.f) - p);
  if (g_ent && p > 0.f) d -= g_ent[b] * p * (logf(p) + entropy[b]);
  dlogits[idx] = d;
}

// ---- dropout ----------------------------------------------------------------------------------
__global__ void dropout_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, float p,
                               const uint64_t* __restrict__ rng, uint64_t call_off, int vec) {
  const int64_t blk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t e0 = blk * 8;
  if (e0 >= n) return;
  const Philox8 r = philox8(rng[0], rng[1] + call_off, (uint64_t)blk);
  const uint32_t thr = drop_threshold(p);
  const float sc = 1.0f / (1.0f - p);
  if (vec && e0 + 8 <= n) {                                 // 16-byte aligned tensors: two float4 each way
    const float4 a = __ldg(reinterpret_cast<const float4*>(x + e0)), b = __ldg(reinterpret_cast<const float4*>(x + e0) + 1);
    float4 oa, ob;
    oa.x = philox_keep(r, 0, thr) ? a.x * sc : 0.f; oa.y = philox_keep(r, 1, thr) ? a.y * sc : 0.f;
    oa.z = philox_keep(r, 2, thr) ? a.z * sc : 0.f; oa.w = philox_keep(r, 3, thr) ? a.w * sc : 0.f;
    ob.x = philox_keep(r, 4, thr) ? b.x * sc : 0.f; ob.y = philox_keep(r, 5, thr) ? b.y * sc : 0.f;
    ob.z = philox_keep(r, 6, thr) ? b.z * sc : 0.f; ob.w = philox_keep(r, 7, thr) ? b.w * sc : 0.f;
    reinterpret_cast<float4*>(y + e0)[0] = oa;
    reinterpret_cast<float4*>(y + e0)[1] = ob;
    return;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (e0 + j < n) y[e0 + j] = philox_keep(r, j, thr) ? x[e0 + j] * sc : 0.f;
}

// nn.Embedding + nn.Dropout of EncoderLSTM (units.py:48-52) in one pass: y[r, :] = drop(emb[tok[r], :]); the keep
// mask is the one vln_dropout / vln_dropout_mask give for the dense [rows, E] tensor under (rng, call_off).
__global__ void embed_drop_fwd_kernel(const int64_t* __restrict__ tok, const float* __restrict__ emb, float* __restrict__ y,
                                      int64_t rows, int E, int V, float p, const uint64_t* __restrict__ rng, uint64_t call_off) {
  const int64_t blk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t e0 = blk * 8, n = rows * E;
  if (e0 >= n) return;
  const Philox8 r = philox8(rng[0], rng[1] + call_off, (uint64_t)blk);
  const uint32_t thr = drop_threshold(p);
  const float sc = 1.0f / (1.0f - p);
  if ((E & 7) == 0) {                                       // the 8 elements share a row; 32-byte aligned both sides
    const int64_t row = e0 / E;
    const int col = (int)(e0 - row * E);
    int64_t t = tok[row];
    t = t < 0 ? 0 : (t >= V ? V - 1 : t);
    const float4* src = reinterpret_cast<const float4*>(emb + t * E + col);
    const float4 a = __ldg(src), b = __ldg(src + 1);
    float4 oa, ob;
    oa.x = philox_keep(r, 0, thr) ? a.x * sc : 0.f; oa.y = philox_keep(r, 1, thr) ? a.y * sc : 0.f;
    oa.z = philox_keep(r, 2, thr) ? a.z * sc : 0.f; oa.w = philox_keep(r, 3, thr) ? a.w * sc : 0.f;
    ob.x = philox_keep(r, 4, thr) ? b.x * sc : 0.f; ob.y = philox_keep(r, 5, thr) ? b.y * sc : 0.f;
    ob.z = philox_keep(r, 6, thr) ? b.z * sc : 0.f; ob.w = philox_keep(r, 7, thr) ? b.w * sc : 0.f;
    reinterpret_cast<float4*>(y + e0)[0] = oa;
    reinterpret_cast<float4*>(y + e0)[1] = ob;
    return;
  }
  for (int j = 0; j < 8 && e0 + j < n; ++j) {
    const int64_t row = (e0 + j) / E;
    const int col = (int)(e0 + j - row * E);
    int64_t t = tok[row];
    t = t < 0 ? 0 : (t >= V ? V - 1 : t);
    y[e0 + j] = philox_keep(r, j, thr) ? __ldg(emb + t * E + col) * sc : 0.f;
  }
}

// Its backward: d_emb[v, :] = sum over the rows r with tok[r] == v of mask(r, :) * d_y[r, :] / (1 - p); the padding
// row gets zero (nn.Embedding(padding_idx=0), units.py:33).  One CTA per vocabulary entry; each of its 8 warps scans
// a contiguous eighth of the token rows (32 per coalesced load + ballot) and accumulates the matching rows in row
// order — a lane owns 8 consecutive columns, i.e. one Philox block per matching row — then the 8 partial sums are
// added in warp order: deterministic, unlike an atomic scatter (ATen's embedding_dense_backward sorts instead).
__global__ void __launch_bounds__(256)
embed_drop_bwd_kernel(const int64_t* __restrict__ tok, const float* __restrict__ dy, float* __restrict__ d_emb, int64_t rows,
                      int E, int padding_idx, float p, const uint64_t* __restrict__ rng, uint64_t call_off) {
  __shared__ float part[8][1024];
  const int v = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t thr = drop_threshold(p);
  const float sc = 1.0f / (1.0f - p);
  const uint64_t seed = rng[0], off = rng[1] + call_off;
  float acc[4][8];                                          // columns [256 c + 8 lane, +8), c < 4 (E <= 1024)
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[c][j] = 0.f;
  if (v != padding_idx) {
    const int64_t per = ((rows + 7) / 8 + 31) / 32 * 32;    // rows per warp, a multiple of 32
    const int64_t r_begin = warp * per, r_end = min(rows, r_begin + per);
    for (int64_t r0 = r_begin; r0 < r_end; r0 += 128) {     // four coalesced token loads in flight
      unsigned m[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t r = r0 + u * 32 + lane;
        m[u] = __ballot_sync(0xffffffffu, r < r_end && __ldg(tok + r) == (int64_t)v);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        unsigned mm = m[u];
        while (mm) {
          const int bit = __ffs(mm) - 1;
          mm &= mm - 1;
          const int64_t r = r0 + u * 32 + bit;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int col = c * 256 + lane * 8;
            if (col >= E) break;
            if ((E & 7) == 0) {                             // the 8 columns are one Philox block of the dense [rows, E] tensor
              const float4 a = __ldg(reinterpret_cast<const float4*>(dy + r * E + col));
              const float4 bq = __ldg(reinterpret_cast<const float4*>(dy + r * E + col) + 1);
              const Philox8 ph = philox8(seed, off, (uint64_t)((r * E + col) >> 3));
              const float x[8] = {a.x, a.y, a.z, a.w, bq.x, bq.y, bq.z, bq.w};
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (philox_keep(ph, j, thr)) acc[c][j] += x[j] * sc;
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (col + j < E) {
                  const int64_t idx = r * E + col + j;
                  if (philox_keep(philox8(seed, off, (uint64_t)(idx >> 3)), (int)(idx & 7), thr)) acc[c][j] += __ldg(dy + idx) * sc;
                }
              }
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) part[warp][c * 256 + lane * 8 + j] = acc[c][j];
  __syncthreads();
  for (int col = tid; col < E; col += 256) {
    float s_ = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s_ += part[w][col];
    d_emb[(size_t)v * E + col] = s_;
  }
}

__global__ void rng_advance_kernel(uint64_t* rng, uint64_t delta) { rng[1] += delta; }

__global__ void dropout_mask_kernel(uint8_t* __restrict__ mask, int64_t n, float p, const uint64_t* __restrict__ rng,
                                    uint64_t call_off) {
  const int64_t blk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t e0 = blk * 8;
  if (e0 >= n) return;
  const Philox8 r = philox8(rng[0], rng[1] + call_off, (uint64_t)blk);
  const uint32_t thr = drop_threshold(p);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (e0 + j < n) mask[e0 + j] = philox_keep(r, j, thr) ? 1 : 0;
}

// Packed keep-bits of the feature dropout for a whole rollout (policy.py:226-231 draws a fresh
// [B,36,2048] mask every decoder step).  Row r of step t is 2048 features = 256 Philox blocks of 8;
// output byte (c / 128) * 128 + (c % 32) * 4 + (c % 128) / 32 of that row holds the 8 keep-bits of block c:
// each half of the row (one CTA of the panorama kernel's cluster) is 128 contiguous bytes, and lane c % 32
// of a warp reads the four bytes of its four 16-byte column chunks at once.
// Same bits as vln_dropout_mask on the dense [rows, 2048] tensor with call_off = off0 + t * off_stride.
// The buffer holds ld_rows rows per step; rows [row0, row0 + rows) of each step are written (numbered row0.. in the stream).
__global__ void feature_mask_bits_kernel(uint8_t* __restrict__ bits, int64_t rows, int n_steps, float p,
                                         const uint64_t* __restrict__ rng, uint64_t off0, uint64_t off_stride,
                                         int64_t ld_rows, int64_t row0) {
  const int64_t total = rows * 256 * n_steps;
  const uint32_t thr = drop_threshold(p);
  const uint64_t seed = rng[0], base = rng[1];
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row_t = o >> 8;                                   // (t, row)
    const int w = (int)(o & 255);                                   // byte position inside the row
    const int c = (w >> 7) * 128 + (w & 3) * 32 + ((w & 127) >> 2);  // Philox block of the row stored there
    const int64_t t = row_t / rows, row = row0 + row_t - t * rows;
    const Philox8 r = philox8(seed, base + off0 + (uint64_t)t * off_stride, (uint64_t)(row * 256 + c));
    uint32_t b = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) b |= (philox_keep(r, j, thr) ? 1u : 0u) << j;
    bits[(t * ld_rows + row) * 256 + w] = (uint8_t)b;
  }
}

// ---- environment on index tables -----------------------------------------------------------------
__device__ __forceinline__ int teacher_slot(int cur, int goal, const int32_t* cand_vp, const int32_t* n_cand,
                                            const int32_t* next_hop, const int64_t* sq_off, const int32_t* vp_local) {
  const int n = n_cand[cur];
  if (cur == goal) return n;                               // STOP (base.py:174-177)
  const int nh = next_hop[sq_off[cur] + vp_local[goal]];
  for (int j = 0; j < n; ++j)
    if (cand_vp[(size_t)cur * VLN_CMAX + j] == nh) return j;
  return n;
}

__global__ void env_observe_kernel(const int32_t* __restrict__ vp, const uint8_t* __restrict__ ended,
                                   const int32_t* __restrict__ goal, const int32_t* __restrict__ cand_vp,
                                   const int32_t* __restrict__ n_cand, const int32_t* __restrict__ next_hop,
                                   const float* __restrict__ dist_tbl, const int64_t* __restrict__ sq_off,
                                   const int32_t* __restrict__ vp_local, int32_t* __restrict__ teacher,
                                   float* __restrict__ dist, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int cur = vp[b], g = goal[b];
  if (teacher) teacher[b] = (ended && ended[b]) ? -1 : teacher_slot(cur, g, cand_vp, n_cand, next_hop, sq_off, vp_local);
  if (dist) dist[b] = dist_tbl[sq_off[cur] + vp_local[g]];
}

__global__ void env_step_kernel(const int32_t* __restrict__ vp_in, const int32_t* __restrict__ view_in,
                                const uint8_t* __restrict__ ended_in, const float* __restrict__ dist_in,
                                const int32_t* __restrict__ goal, const int32_t* __restrict__ action,
                                const int32_t* __restrict__ cand_vp, const int32_t* __restrict__ cand_view,
                                const int32_t* __restrict__ n_cand, const int32_t* __restrict__ next_hop,
                                const float* __restrict__ dist_tbl, const int64_t* __restrict__ sq_off,
                                const int32_t* __restrict__ vp_local, int32_t* __restrict__ vp_out,
                                int32_t* __restrict__ view_out, uint8_t* __restrict__ ended_out,
                                float* __restrict__ dist_out, int32_t* __restrict__ teacher,
                                float* __restrict__ reward, float* __restrict__ mask, int32_t* __restrict__ n_active,
                                int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  bool running = false;
  if (b < B) {
    int cur = vp_in[b], vw = view_in[b];
    const int g = goal[b];
    const bool was_ended = ended_in[b] != 0;
    const int a = action[b];
    const bool stop = was_ended || a < 0 || a >= n_cand[cur];       // envdrop.py:198-203
    if (!stop) {
      vw = cand_view[(size_t)cur * VLN_CMAX + a];                   // agent turns to the view, then steps (misc.py:366-390)
      cur = cand_vp[(size_t)cur * VLN_CMAX + a];
    }
    vp_out[b] = cur;
    view_out[b] = vw;
    const float d = dist_tbl[sq_off[cur] + vp_local[g]];
    if (reward) {                                                   // envdrop.py:207-213
      float r = 0.f;
      if (!was_ended) {
        if (stop) r = d < 3.0f ? 2.f : -2.f;
        else { const float dd = dist_in[b] - d; r = dd > 0.f ? 1.f : (dd < 0.f ? -1.f : 0.f); }
      }
      reward[b] = r;
    }
    if (mask) mask[b] = was_ended ? 0.f : 1.f;
    dist_out[b] = d;
    const bool now_ended = was_ended || stop;
    ended_out[b] = now_ended ? 1 : 0;
    running = !now_ended;
    if (teacher) teacher[b] = now_ended ? -1 : teacher_slot(cur, g, cand_vp, n_cand, next_hop, sq_off, vp_local);
  }
  if (n_active) {
    const unsigned m = __ballot_sync(0xffffffffu, running);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_active, __popc(m));
  }
}

// ---- A2C loss assembly (envdrop.py:240-264): one thread per episode, backwards in time ---------------
__global__ void a2c_fwd_kernel(const float* __restrict__ reward, const float* __restrict__ mask,
                               const float* __restrict__ logp, const float* __restrict__ entropy,
                               const float* __restrict__ value, const float* __restrict__ last_value,
                               const uint8_t* __restrict__ ended, float gamma, float ent_coef,
                               float* __restrict__ loss_b, float* __restrict__ ret, float* __restrict__ total,
                               float* __restrict__ critic_sq, int T, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  float tot = 0.f, csq = 0.f;
  if (b < B) {
    const float lv = ended[b] ? 0.f : last_value[b];               // (~ended) * last_value, float32
    double R = 0.0;
    float acc = 0.f;
    for (int t = T - 1; t >= 0; --t) {
      const size_t i = (size_t)t * B + b;
      // numpy promotion: the first product is float32 * python float, every later one float64
      R = (t == T - 1 ? (double)(lv * gamma) : R * (double)gamma) + (double)reward[i];
      const double m = mask[i];
      const double adv = R - (double)value[i];
      float cur = (float)(-(double)logp[i] * adv * m);
      cur = (float)((double)cur + adv * adv * m * 0.5);
      if (ent_coef != 0.f) cur = (float)((double)cur + (double)(-ent_coef * entropy[i]) * m);
      acc += cur;
      ret[i] = (float)R;
      tot += mask[i];
      csq += (float)(adv * adv * m);
    }
    loss_b[b] = acc;
  }
  tot = warp_sum(tot);
  csq = warp_sum(csq);
  if ((threadIdx.x & 31) == 0) {
    if (total) atomicAdd(total, tot);
    if (critic_sq) atomicAdd(critic_sq, csq);
  }
}

__global__ void a2c_bwd_kernel(const float* __restrict__ g_b, const float* __restrict__ mask,
                               const float* __restrict__ value, const float* __restrict__ ret, float ent_coef,
                               float* __restrict__ d_logp, float* __restrict__ d_value,
                               float* __restrict__ d_entropy, int T, int B) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= T * B) return;
  const int b = idx % B;
  const float g = g_b[b] * mask[idx];
  const float adv = ret[idx] - value[idx];
  d_logp[idx] = -g * adv;
  d_value[idx] = -g * adv;
  if (d_entropy) d_entropy[idx] = -g * ent_coef;
}

// ---- gradient clip + optimiser ----------------------------------------------------------------------
struct Groups {
  int64_t off[5];
  float max_norm[4];
  int n;
};

// Group offsets are multiples of 4 floats (FlatOptimizer pads every group), so a float4 never straddles groups.
__device__ __forceinline__ int group_of(const Groups& gr, int64_t i) {
  int k = 0;
  while (k + 1 < gr.n && i >= gr.off[k + 1]) ++k;
  return k;
}

// Per-group sum of squares, DETERMINISTIC: every block writes its partial sums to a scratch slot, the last block to
// arrive (ticket) adds the slots in index order.  (A plain atomicAdd per block gives an order-dependent fp32 sum:
// under data parallelism the ranks then compute clip coefficients that differ in the last bit and their parameters
// drift apart by an ulp per step — found by tools/dp_check.py.)
constexpr int kSqBlocks = 296;
__device__ float g_sq_partials[kSqBlocks * 4];
__device__ unsigned int g_sq_ticket;

__global__ void sqnorm_kernel(const float* __restrict__ grad, Groups gr, float scale, float* __restrict__ sqnorm) {
  __shared__ float red[4][32];
  __shared__ bool last;
  const int64_t n4 = gr.off[gr.n] >> 2;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(grad) + i);
    const float s = (g.x * scale) * (g.x * scale) + (g.y * scale) * (g.y * scale) + (g.z * scale) * (g.z * scale) +
                    (g.w * scale) * (g.w * scale);
    const int k = group_of(gr, i << 2);
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] += (j == k) ? s : 0.f;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float s = warp_sum(acc[k]);
    if (lane == 0) red[k][warp] = s;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float s = lane < (blockDim.x >> 5) ? red[k][lane] : 0.f;
      s = warp_sum(s);
      if (lane == 0) g_sq_partials[blockIdx.x * 4 + k] = s;
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(&g_sq_ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (threadIdx.x < 4) {                                    // fixed order: block 0, 1, 2, ...
    float s = 0.f;
    for (unsigned int b = 0; b < gridDim.x; ++b) s += __ldcg(&g_sq_partials[b * 4 + threadIdx.x]);
    if ((int)threadIdx.x < gr.n) sqnorm[threadIdx.x] = s;
    if (threadIdx.x == 0) g_sq_ticket = 0;                  // ready for the next launch
  }
}

__device__ __forceinline__ float rmsprop1(float& sq, float g, float p, float lr) {
  sq = 0.99f * sq + 0.01f * g * g;
  return p - lr * g / (sqrtf(sq) + 1e-8f);
}
__device__ __forceinline__ float adam1(float& m, float& v, float g, float p, float lr1, float rbc2) {
  m = 0.9f * m + 0.1f * g;
  v = 0.999f * v + 0.001f * g * g;
  return p - lr1 * m / (sqrtf(v) * rbc2 + 1e-8f);
}

__global__ void optim_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ s1,
                             float* __restrict__ s2, Groups gr, const float* __restrict__ sqnorm, float scale,
                             int kind, float lr, float bc1, float bc2) {
  const int64_t n4 = gr.off[gr.n] >> 2;
  float coef[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {                                   // clip_grad_norm_: coef = max/(norm+1e-6), clamped to 1
    coef[k] = scale;
    if (k < gr.n && gr.max_norm[k] > 0.f) {
      const float c = gr.max_norm[k] / (sqrtf(sqnorm[k]) + 1e-6f);
      if (c < 1.f) coef[k] = scale * c;
    }
  }
  const float lr1 = lr / bc1, rbc2 = 1.0f / sqrtf(bc2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = group_of(gr, i << 2);
    const float c = k == 0 ? coef[0] : (k == 1 ? coef[1] : (k == 2 ? coef[2] : coef[3]));
    float4 g = __ldg(reinterpret_cast<const float4*>(grad) + i);
    g.x *= c; g.y *= c; g.z *= c; g.w *= c;
    float4 p = reinterpret_cast<float4*>(param)[i];
    float4 a = reinterpret_cast<float4*>(s1)[i];
    if (kind == 0) {                                              // torch.optim.RMSprop defaults
      p.x = rmsprop1(a.x, g.x, p.x, lr); p.y = rmsprop1(a.y, g.y, p.y, lr);
      p.z = rmsprop1(a.z, g.z, p.z, lr); p.w = rmsprop1(a.w, g.w, p.w, lr);
    } else {                                                      // torch.optim.Adam defaults
      float4 v = reinterpret_cast<float4*>(s2)[i];
      p.x = adam1(a.x, v.x, g.x, p.x, lr1, rbc2); p.y = adam1(a.y, v.y, g.y, p.y, lr1, rbc2);
      p.z = adam1(a.z, v.z, g.z, p.z, lr1, rbc2); p.w = adam1(a.w, v.w, g.w, p.w, lr1, rbc2);
      reinterpret_cast<float4*>(s2)[i] = v;
    }
    reinterpret_cast<float4*>(s1)[i] = a;
    reinterpret_cast<float4*>(param)[i] = p;
  }
}

}  // namespace

#define STREAM ((cudaStream_t)stream)

extern "C" int vln_lstm_pointwise_fwd(const float* gates, const float* c0, float* h1, float* c1, float* acts, int B,
                                      int H, void* stream) {
  VLN_REQUIRE(gates && c0 && h1 && c1 && B > 0 && H > 0, "bad arguments");
  lstm_pw_fwd_kernel<<<(B * H + 255) / 256, 256, 0, STREAM>>>(gates, c0, h1, c1, acts, B, H);
  VLN_LAUNCH_OK();
  return 0;
}

extern "C" int vln_lstm_pointwise_bwd(const float* acts, const float* c0, const float* c1, const float* d_h1,
                                      const float* d_c1, float* d_gates, float* d_c0, int B, int H, void* stream) {
  VLN_REQUIRE(acts && c0 && c1 && d_gates && d_c0 && B > 0 && H > 0, "bad arguments");
  lstm_pw_bwd_kernel<<<(B * H + 255) / 256, 256, 0, STREAM>>>(acts, c0, c1, d_h1, d_c1, d_gates, d_c0, B, H);
  VLN_LAUNCH_OK();
  return 0;
}

extern "C" int vln_policy_fwd(const float* logits, const int32_t* target, int feedback, const uint64_t* rng,
                              uint64_t call_off, float* ce, int32_t* action, float* logp, float* entropy, float* probs,
                              int B, void* stream) {
  VLN_REQUIRE(logits && B > 0 && feedback >= 0 && feedback <= 2, "bad arguments");
  VLN_REQUIRE(feedback != 0 || target, "teacher feedback needs targets");
  VLN_REQUIRE(feedback != 2 || rng, "sampling needs an rng state");
  policy_fwd_kernel<<<(B + 3) / 4, 128, 0, STREAM>>>(logits, target, feedback, rng, call_off, ce, action, logp, entropy,
                                                     probs, B);
  VLN_LAUNCH_OK();
  return 0;
}

extern "C" int vln_policy_bwd(const float* probs, const int32_t* target, const int32_t* action, const float* entropy,
                              const float* g_ce, const float* g_logp, const float* g_ent, float* dlogits, int B,
                              void* stream) {
  VLN_REQUIRE(probs && dlogits && B > 0, "bad arguments");
  VLN_REQUIRE(!g_ce || target, "g_ce needs targets");
  VLN_REQUIRE(!g_logp || action, "g_logp needs actions");
  VLN_REQUIRE(!g_ent || entropy, "g_ent needs entropy");
  policy_bwd_kernel<<<(B * VLN_NSLOT + 255) / 256, 256, 0, STREAM>>>(probs, target, action, entropy, g_ce, g_logp,
                                                                     g_ent, dlogits, B);
  VLN_LAUNCH_OK();
  return 0;
}

extern "C" int vln_rng_advance(uint64_t* rng, uint64_t delta, void* stream) {
  VLN_REQUIRE(rng, "bad arguments");
  rng_advance_kernel<<<1, 1, 0, STREAM>>>(rng, delta);
  VLN_LAUNCH_OK();
  return 0;
}

extern "C" int vln_dropout(const float* x, float* y, int64_t n, float p, const uint64_t* rng, uint64_t call_off,
                           void* stream) {
  VLN_REQUIRE(x && y && rng && n > 0 && p >= 0.f && p < 1.f, "bad arguments");
  const int64_t blks = (n + 7) / 8;
  const int vec = (((uintptr_t)x | (uintptr_t)y) & 15) == 0;
  dropout_kernel<<<(unsigned)((blks + 255) / 256), 256, 0, STREAM>>>(x, y, n, p, rng, call_off, vec);
  VLN_LAUNCH_OK();
  return 0;
}

extern "C" int vln_embed_drop_fwd(const int64_t* tokens, const float* emb, float* y, int64_t rows, int E, int V, float p,
                                  const uint64_t* rng, uint64_t call_off, void* stream) {
  VLN_REQUIRE(tokens && emb && y && rng && rows > 0 && E > 0 && V > 0 && p >= 0.f && p < 1.f, "bad arguments");
  VLN_REQUIRE((((uintptr_t)emb | (uintptr_t)y) & 15) == 0, "embedding table and output must be 16-byte aligned");
  const int64_t blks = (rows * E + 7) / 8;
  embed_drop_fwd_kernel<<<(unsigned)((blks + 255) / 256), 256, 0, STREAM>>>(tokens, emb, y, rows, E, V, p, rng, call_off);
  VLN_LAUNCH_OK();
  return 0;
}

extern "C" int vln_embed_drop_bwd(const int64_t* tokens, const float* d_y, float* d_emb, int64_t rows, int E, int V,
                                  int padding_idx, float p, const uint64_t* rng, uint64_t call_off, void* stream) {
  VLN_REQUIRE(tokens && d_y && d_emb && rng && rows > 0 && V > 0 && p >= 0.f && p < 1.f, "bad arguments");
  VLN_REQUIRE(E > 0 && E <= 1024, "embedding width must be at most 1024");
  embed_drop_bwd_kernel<<<V, 256, 0, STREAM>>>(tokens, d_y, d_emb, rows, E, padding_idx, p, rng, call_off);
  VLN_LAUNCH_OK();
  return 0;
}

extern "C" int vln_dropout_mask(uint8_t* mask, int64_t n, float p, const uint64_t* rng, uint64_t call_off,
                                void* stream) {
  VLN_REQUIRE(mask && rng && n > 0 && p >= 0.f && p < 1.f, "bad arguments");
  const int64_t blks = (n + 7) / 8;
  dropout_mask_kernel<<<(unsigned)((blks + 255) / 256), 256, 0, STREAM>>>(mask, n, p, rng, call_off);
  VLN_LAUNCH_OK();
  return 0;
}

extern "C" int vln_feature_mask_bits(uint8_t* bits, int64_t rows, int n_steps, float p, const uint64_t* rng,
                                     uint64_t off0, uint64_t off_stride, void* stream) {
  VLN_REQUIRE(bits && rng && rows > 0 && n_steps > 0 && p > 0.f && p < 1.f, "bad arguments");
  const int64_t total = rows * 256 * n_steps;
  const int64_t want = (total + 255) / 256;
  const unsigned grid = (unsigned)(want < 148 * 16 ? want : 148 * 16);
  feature_mask_bits_kernel<<<grid, 256, 0, STREAM>>>(bits, rows, n_steps, p, rng, off0, off_stride, rows, 0);
  VLN_LAUNCH_OK();
  return 0;
}

// The same bits for a sub-block of a larger buffer: steps [0, n_steps), rows [row0, row0 + rows) of a buffer with
// ld_rows rows per step (paired rollouts: the teacher-forced half only needs the first T_teacher steps).
extern "C" int vln_feature_mask_bits_ld(uint8_t* bits, int64_t rows, int64_t ld_rows, int64_t row0, int n_steps, float p,
                                        const uint64_t* rng, uint64_t off0, uint64_t off_stride, void* stream) {
  VLN_REQUIRE(bits && rng && rows > 0 && n_steps > 0 && p > 0.f && p < 1.f && row0 >= 0 && row0 + rows <= ld_rows,
              "bad arguments");
  const int64_t total = rows * 256 * n_steps;
  const int64_t want = (total + 255) / 256;
  const unsigned grid = (unsigned)(want < 148 * 16 ? want : 148 * 16);
  feature_mask_bits_kernel<<<grid, 256, 0, STREAM>>>(bits, rows, n_steps, p, rng, off0, off_stride, ld_rows, row0);
  VLN_LAUNCH_OK();
  return 0;
}

extern "C" int vln_env_observe(const int32_t* vp, const uint8_t* ended, const int32_t* goal, const int32_t* cand_vp,
                               const int32_t* n_cand, const int32_t* next_hop, const float* dist_tbl,
                               const int64_t* sq_off, const int32_t* vp_local, int32_t* teacher, float* dist, int B,
                               void* stream) {
  VLN_REQUIRE(vp && goal && cand_vp && n_cand && next_hop && dist_tbl && sq_off && vp_local && B > 0, "bad arguments");
  env_observe_kernel<<<(B + 127) / 128, 128, 0, STREAM>>>(vp, ended, goal, cand_vp, n_cand, next_hop, dist_tbl, sq_off,
                                                          vp_local, teacher, dist, B);
  VLN_LAUNCH_OK();
  return 0;
}

extern "C" int vln_env_step(const int32_t* vp_in, const int32_t* view_in, const uint8_t* ended_in,
                            const float* dist_in, const int32_t* goal, const int32_t* action,
                            const int32_t* cand_vp, const int32_t* cand_view, const int32_t* n_cand,
                            const int32_t* next_hop, const float* dist_tbl, const int64_t* sq_off,
                            const int32_t* vp_local, int32_t* vp_out, int32_t* view_out, uint8_t* ended_out,
                            float* dist_out, int32_t* teacher, float* reward, float* mask, int32_t* n_active, int B,
                            void* stream) {
  VLN_REQUIRE(vp_in && view_in && ended_in && dist_in && goal && action && cand_vp && cand_view && n_cand &&
                  next_hop && dist_tbl && sq_off && vp_local && vp_out && view_out && ended_out && dist_out && B > 0,
              "bad arguments");
  env_step_kernel<<<(B + 127) / 128, 128, 0, STREAM>>>(vp_in, view_in, ended_in, dist_in, goal, action, cand_vp,
                                                       cand_view, n_cand, next_hop, dist_tbl, sq_off, vp_local, vp_out,
                                                       view_out, ended_out, dist_out, teacher, reward, mask, n_active,
                                                       B);
  VLN_LAUNCH_OK();
  return 0;
}

extern "C" int vln_a2c_fwd(const float* reward, const float* mask, const float* logp, const float* entropy,
                           const float* value, const float* last_value, const uint8_t* ended, float gamma,
                           float ent_coef, float* loss_b, float* ret, float* total, float* critic_sq, int T, int B,
                           void* stream) {
  VLN_REQUIRE(reward && mask && logp && value && last_value && ended && loss_b && ret && T > 0 && B > 0,
              "bad arguments");
  VLN_REQUIRE(ent_coef == 0.f || entropy, "entropy term needs the entropies");
  a2c_fwd_kernel<<<(B + 127) / 128, 128, 0, STREAM>>>(reward, mask, logp, entropy, value, last_value, ended, gamma,
                                                      ent_coef, loss_b, ret, total, critic_sq, T, B);
  VLN_LAUNCH_OK();
  return 0;
}

extern "C" int vln_a2c_bwd(const float* g_b, const float* mask, const float* value, const float* ret,
                           float ent_coef, float* d_logp, float* d_value, float* d_entropy, int T, int B,
                           void* stream) {
  VLN_REQUIRE(g_b && mask && value && ret && d_logp && d_value && T > 0 && B > 0, "bad arguments");
  a2c_bwd_kernel<<<(T * B + 255) / 256, 256, 0, STREAM>>>(g_b, mask, value, ret, ent_coef, d_logp, d_value,
                                                          d_entropy, T, B);
  VLN_LAUNCH_OK();
  return 0;
}

static int make_groups(Groups* g, const int64_t* group_off, const float* max_norm, int n_groups) {
  if (n_groups < 1 || n_groups > 4) return -1;
  g->n = n_groups;
  for (int i = 0; i <= n_groups; ++i) g->off[i] = group_off[i];
  for (int i = 0; i < n_groups; ++i) g->max_norm[i] = max_norm ? max_norm[i] : 0.f;
  return 0;
}

extern "C" int vln_grad_sqnorm(const float* grad, const int64_t* group_off, int n_groups, float* sqnorm,
                               float grad_scale, void* stream) {
  VLN_REQUIRE(grad && group_off && sqnorm, "bad arguments");
  Groups g;
  VLN_REQUIRE(make_groups(&g, group_off, nullptr, n_groups) == 0, "1..4 groups supported");
  for (int i = 0; i <= n_groups; ++i) VLN_REQUIRE(group_off[i] % 4 == 0, "group offsets must be multiples of 4 floats");
  VLN_REQUIRE(((uintptr_t)grad & 15) == 0, "grad must be 16-byte aligned");
  sqnorm_kernel<<<kSqBlocks, 256, 0, STREAM>>>(grad, g, grad_scale, sqnorm);
  VLN_LAUNCH_OK();
  return 0;
}

extern "C" int vln_optim_step(float* param, const float* grad, float* state1, float* state2, const int64_t* group_off,
                              const float* max_norm, int n_groups, const float* sqnorm, float grad_scale, int kind,
                              float lr, int step, void* stream) {
  VLN_REQUIRE(param && grad && state1 && group_off && sqnorm && (kind == 0 || (kind == 1 && state2)), "bad arguments");
  Groups g;
  VLN_REQUIRE(make_groups(&g, group_off, max_norm, n_groups) == 0, "1..4 groups supported");
  for (int i = 0; i <= n_groups; ++i) VLN_REQUIRE(group_off[i] % 4 == 0, "group offsets must be multiples of 4 floats");
  VLN_REQUIRE((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)state1 | (uintptr_t)state2) & 15) == 0, "buffers must be 16-byte aligned");
  const float bc1 = 1.0f - powf(0.9f, (float)step), bc2 = 1.0f - powf(0.999f, (float)step);
  optim_kernel<<<592, 256, 0, STREAM>>>(param, grad, state1, state2, g, sqnorm, grad_scale, kind, lr, bc1, bc2);
  VLN_LAUNCH_OK();
  return 0;
}
