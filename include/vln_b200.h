/*
 * vln_b200.h — C-ABI of the B200-native rollout hot path.
 *
 * The reference (IMNearth/Curriculum-Learning-For-VLN, tasks/R2R-judy/src) is pure
 * Python/PyTorch and has no FFI of its own; its boundary for this path is the set of
 * nn.Module.forward()/agent helper calls listed in SURVEY.md §8(a,b).  Each entry point
 * below names the reference code it replaces.  Host code (Python, ctypes) sits on top and
 * keeps the reference's module/agent signatures.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (a torch tensor's data_ptr())
 *     unless it says "host"; the library allocates nothing persistent except vln_ctx;
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue, never synchronise;
 *   - return 0 on success, a negative code on failure; vln_last_error() gives the text
 *     (thread-local);  nothing throws or exits;
 *   - layouts are row-major, innermost dimension last.  F = 2176 = 2048 image + 128 angle,
 *     V = 36 views, CMAX = 15 neighbours (+1 END slot => 16 action slots).
 *   - randomness: `rng` is a DEVICE pointer to two uint64 {seed, base}; a call site passes its
 *     own `call_off` and the kernel uses the stream (seed, base + call_off).  Keeping the base in
 *     device memory lets a captured CUDA graph draw fresh masks on every replay (vln_rng_advance).
 *   - dropout: keep-mask bit for element e of a tensor = Philox4x32-10(seed, offset, e/8)
 *     16-bit lane (e%8) >= round(p*65536); kept values are scaled by 1/(1-p) (nn.Dropout).
 *     vln_dropout_mask() exposes the very same mask so the oracle can be fed identical masks.
 */
#ifndef VLN_B200_H_
#define VLN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VLN_V 36
#define VLN_IMG 2048
#define VLN_ANG 128
#define VLN_F 2176
#define VLN_CMAX 15
#define VLN_NSLOT 16

typedef struct vln_ctx vln_ctx;

int vln_version(void);
const char* vln_last_error(void);

/* Feature store.  Replaces ImageFeatures.read_in's dict + EnvBatch.getStates lookup
 * (misc.py:254-279, common_env.py:72-89): one HBM-resident bf16 table [n_vp,36,2048]
 * described to the TMA unit as a 2-D tensor [n_vp*36, 2048]. */
int vln_ctx_create(vln_ctx** out, const void* table_bf16, int n_vp, int device);
void vln_ctx_destroy(vln_ctx* ctx);

/* R2RBatch.observe feature concat + _feature_variable (common_env.py:309, base.py:141-147):
 * out[b] = concat(table[vp[b]], loc4[view[b]] repeated x32) as fp32 [B,36,2176]. bit-exact. */
int vln_gather_pano(const vln_ctx* ctx, const int32_t* vp, const int32_t* view,
                    const float* loc4 /*[36,36,4]*/, float* out, int B, void* stream);

/* make_candidate feature + _candidate_variable (common_env.py:283-291, base.py:149-157):
 * out[b,j] = concat(table[vp[b], cand_view[vp[b],j]], cand_ang4[vp[b],j,view[b]%12] x32) for
 * j < n_cand[vp[b]], zeros elsewhere (row n_cand = END).  fp32 [B,C,2176]; out_len[b]=n_cand+1. */
int vln_gather_cand(const vln_ctx* ctx, const int32_t* vp, const int32_t* view,
                    const int32_t* cand_view /*[n_vp,15]*/, const float* cand_ang4 /*[n_vp,15,12,4]*/,
                    const int32_t* n_cand /*[n_vp]*/, float* out, int32_t* out_len, int B, int C,
                    void* stream);

/* Fused gather + soft-dot attention over the 36-view panorama (SoftDotAttention.forward,
 * units.py:107-118, with EnvDropDecoder's in-place feature dropout policy.py:226-231 folded
 * into the load).  Each table row is read from HBM once; [B,36,2176] is never materialised.
 *   mode 0 (forward):  vec = query q[B,2176];  attn_io <- softmax_v(x~_v . q) [B,36];
 *                      out <- sum_v attn_v x~_v  [B,2176];  fwd_out unused (may be NULL)
 *   mode 1 (backward): vec = d(out) [B,2176], attn_io = saved attn (read), fwd_out = saved forward
 *                      `out`;  out <- dq = sum_v dlogit_v x~_v,  dlogit = attn*(r - attn.r), r_v = x~_v . vec
 * x~ = dropout(table row) (+) angle embedding.  split (historical name) selects the kernel: 1 = automatic
 * (by B), 2 = low-latency kernel (one episode per 2-CTA cluster, csrc/pano_attn.cu), 4 = streaming kernel
 * (4-row chunks through a ring, single-pass online softmax, csrc/pano_stream.cu; needs drop_p == 0 or
 * mask_bits, otherwise the call falls back to the cluster kernel). */
#define VLN_SPLIT_MAX_B 1024
int vln_pano_attn(const vln_ctx* ctx, const int32_t* vp, const int32_t* view, const float* loc4,
                  const float* vec, float* attn_io, const float* fwd_out, float* out, int B, int mode,
                  float drop_p, const uint64_t* rng, uint64_t call_off, int split, void* stream);

/* Same kernel with explicit row strides (in floats, multiples of 4, >= 2176) for vec, fwd_out and out, so the
 * result can land inside a wider operand row (the LSTMCell input [act_emb | visual | h] of policy.py:236-238)
 * and the backward can read its d(out) / saved forward output from there. */
/* mask_bits (nullable): pre-generated keep-bits [B,36,256 bytes] of this step's feature dropout
 * (vln_feature_mask_bits) — the kernel then streams 256 mask bytes next to each 4 096-byte row instead of
 * running Philox inline (which made the kernel ALU-bound: 15.4 us vs 7.6 us per launch at B=64).  * mode + 2 (row-strided entry point only): promises that vp / view / mask_bits (and, backward, attn_io) were
 * complete before the PRECEDING kernel of the stream started — true throughout the backward pass, whose indices
 * date from the forward pass — so the first rows are requested before the programmatic-dependency wait. */
int vln_pano_attn_ld(const vln_ctx* ctx, const int32_t* vp, const int32_t* view, const float* loc4,
                     const float* vec, int ld_vec, float* attn_io, const float* fwd_out, int ld_fwd,
                     float* out, int ld_out, int B, int mode, float drop_p, const uint64_t* rng,
                     uint64_t call_off, const uint8_t* mask_bits, int split, void* stream);
/* Packed keep-bits of the feature dropout (policy.py:226-231) for n_steps decoder steps at once:
 * bits [n_steps, rows, 256] bytes, rows = B*36 panorama rows of 2048 features; step t uses the stream
 * (seed, base + off0 + t*off_stride); the bits equal vln_dropout_mask on the dense [rows,2048] tensor.
 * Byte (c/128)*128 + (c%32)*4 + (c%128)/32 of a row = keep-bits of features [8c, 8c+8). */
int vln_feature_mask_bits(uint8_t* bits, int64_t rows, int n_steps, float p, const uint64_t* rng,
                          uint64_t off0, uint64_t off_stride, void* stream);
/* The same bits for rows [row0, row0 + rows) of the first n_steps steps of a buffer that holds ld_rows rows per
 * step (paired rollouts: the teacher-forced half of the batch only lives for the first T_teacher steps). */
int vln_feature_mask_bits_ld(uint8_t* bits, int64_t rows, int64_t ld_rows, int64_t row0, int n_steps, float p,
                             const uint64_t* rng, uint64_t off0, uint64_t off_stride, void* stream);

/* Candidate logits (EnvDropDecoder.candidate_attn policy.py:199-206; also ActionScoring
 * units.py:173-185 after folding its Linear layers into tgt/bias on the host side):
 *   logits[b,j] = x~c[b,j] . tgt[b] + bias[b]   j <= n_cand (END row is all-zero features)
 *   logits[b,j] = -inf                          j >  n_cand          (length2mask, misc.py:481-486)
 * logits is [B,16].  bias may be NULL. */
int vln_cand_logits_fwd(const vln_ctx* ctx, const int32_t* vp, const int32_t* view,
                        const int32_t* cand_view, const float* cand_ang4, const int32_t* n_cand,
                        const float* tgt, const float* bias, float* logits, int B, float drop_p,
                        const uint64_t* rng, uint64_t call_off, void* stream);
/* d_tgt[b] = sum_j dlogits[b,j] x~c[b,j];  d_bias[b] = sum_{j<=n_cand} dlogits[b,j] (may be NULL). */
int vln_cand_logits_bwd(const vln_ctx* ctx, const int32_t* vp, const int32_t* view,
                        const int32_t* cand_view, const float* cand_ang4, const int32_t* n_cand,
                        const float* dlogits, float* d_tgt, float* d_bias, int B, float drop_p,
                        const uint64_t* rng, uint64_t call_off, void* stream);

/* Soft-dot attention over the instruction context (SoftDotAttention.forward units.py:107-118,
 * the bmm/softmax/bmm part; linear_in/linear_out are GEMMs outside).  context fp32 [B,L,H],
 * tgt [B,H], lengths[b] = number of unmasked rows (mask = position >= length). */
int vln_ctx_attn_fwd(const float* context, const float* tgt, const int32_t* lengths, float* attn,
                     float* weighted, int B, int L, int H, void* stream);
/* d_tgt <- sum_l dlogit_l ctx_l ; d_context += attn_l*d_weighted + dlogit_l*tgt  (accumulates);
 * d_attn_ext (nullable) = extra gradient arriving on the attention weights themselves;
 * d_context may be NULL when the context needs no gradient (a materialised feature tensor). */
int vln_ctx_attn_bwd(const float* context, const float* tgt, const int32_t* lengths,
                     const float* attn, const float* d_weighted, const float* d_attn_ext,
                     float* d_tgt, float* d_context, int B, int L, int H, void* stream);

/* Row-strided variants: `weighted` (forward) / `d_weighted` (backward) rows are ld floats apart, so the
 * weighted context is written straight into cat((weighted, h)) (units.py:119) and its gradient read from there. */
/* ctx_ready = 1 promises that `context` and `lengths` (backward: also `attn` and `tgt`) were complete before the
 * PRECEDING kernel of the stream started (every decoder step but the first): the tile copy is then requested before
 * the programmatic-dependency wait and overlaps the predecessor's tail. */
int vln_ctx_attn_fwd_ld(const float* context, const float* tgt, const int32_t* lengths, float* attn,
                        float* weighted, int ld_weighted, int B, int L, int H, int ctx_ready, void* stream);
/* dlogit_out (nullable) [B,L] receives dlogit, so that d_context = sum over steps of
 * attn^T d_weighted + dlogit^T tgt can be ONE batched GEMM at the end of a rollout instead of a
 * read-modify-write of the whole context gradient per step (then pass d_context = NULL). */
int vln_ctx_attn_bwd_ld(const float* context, const float* tgt, const int32_t* lengths,
                        const float* attn, const float* d_weighted, int ld_d_weighted, const float* d_attn_ext,
                        float* d_tgt, float* d_context, float* dlogit_out, int B, int L, int H, int ctx_ready,
                        void* stream);

/* nn.LSTMCell pointwise half (policy.py:53,159,238): gates [B,4H] (i,f,g,o pre-activations,
 * biases already added) + c0 -> h1, c1; acts [B,4H] keeps the activated gates for backward. */
int vln_lstm_pointwise_fwd(const float* gates, const float* c0, float* h1, float* c1, float* acts,
                           int B, int H, void* stream);
int vln_lstm_pointwise_bwd(const float* acts, const float* c0, const float* c1, const float* d_h1,
                           const float* d_c1, float* d_gates, float* d_c0, int B, int H, void* stream);

/* LSTMCell pointwise half fused with the nn.Dropout of h_1 that follows it in EnvDropDecoder.forward
 * (policy.py:238-240): h1_drop (nullable; rows ld_drop floats apart) = dropout(h1; p, call_off), written
 * into the [weighted | h] operand of text_attn.linear_out.  Backward: d_h1 = dropout'(d_h1_drop) +
 * d_h1_extra (nullable: the critic's gradient on the undropped h_1, envdrop.py:247) and d_c1 (nullable). */
int vln_lstm_pointwise_drop_fwd(const float* gates, const float* c0, float* h1, float* c1, float* acts,
                                float* h1_drop, int ld_drop, int B, int H, float p, const uint64_t* rng,
                                uint64_t call_off, void* stream);
int vln_lstm_pointwise_drop_bwd(const float* acts, const float* c0, const float* c1, const float* d_h1_drop,
                                int ld_drop, const float* d_h1_extra, const float* d_c1, float* d_gates,
                                float* d_c0, int B, int H, float p, const uint64_t* rng, uint64_t call_off,
                                void* stream);

/* Text-attention stage of a fused EnvDrop decoder step as ONE launch each way (policy.py:238-241 + units.py:107-118):
 * nn.LSTMCell pointwise half (gates [B,4H] already hold the gate GEMM + biases) -> h_1, c_1 (+ saved activations),
 * drop(h_1) -> wh[b, H:2H), then SoftDotAttention's masked softmax over the instruction and the weighted context
 * -> wh[b, 0:H).  The query projection linear_in is folded into the context once per rollout: `cw` = ctx W_in
 * [B,L,H] (logit_l = ctx_l . (W_in h) = cw_l . h), so no per-step GEMM sits between the LSTM cell and the attention.
 * H must be 512 (one thread per hidden unit), L <= 80.  tiles_ready: ctx / cw / lengths were complete before the
 * PRECEDING kernel started (both tiles are then requested ahead of the programmatic-dependency wait).
 * Backward: dwh[b, 0:H) = d_weighted and dwh[b, H:2H) = the linear_out part of d drop(h_1) (both from the
 * linear_out input-gradient GEMM); emits dlogit [B,L] (for d_ctx / d_cw, one batched GEMM per rollout), adds
 * sum_l dlogit_l cw_l to d drop(h_1), undoes the dropout, adds d_h1_extra (nullable: the critic's gradient on h_1,
 * envdrop.py:247) and runs the LSTMCell pointwise backward (d_c1 nullable) -> d_gates [B,4H], d_c0 [B,H]. */
int vln_envdrop_ctx_step_fwd(const float* gates, const float* c0, float* h1, float* c1, float* acts, float* wh,
                             int ld_wh, const float* ctx, const float* cw, const int32_t* lengths, float* attn,
                             int B, int L, int H, float p, const uint64_t* rng, uint64_t call_off, int tiles_ready,
                             void* stream);
int vln_envdrop_ctx_step_bwd(const float* ctx, const float* cw, const int32_t* lengths, const float* attn,
                             const float* dwh, int ld_dwh, float* dlogit_out, const float* acts, const float* c0,
                             const float* c1, const float* d_h1_extra, const float* d_c1, float* d_gates,
                             float* d_c0, int B, int L, int H, float p, const uint64_t* rng, uint64_t call_off,
                             void* stream);

/* vln_linear_bf16x3 (accumulating into a zero-filled y [M,N], M <= 128, N = H) with the per-element glue that follows
 * it in the EnvDrop decoder step folded in as a tile epilogue — one launch less on the step's latency chain each:
 *   state_fwd: y = linear_out pre-activation (units.py:119-120) -> exactly vln_envdrop_state_fwd(y, apply_tanh = 1, ...);
 *   state_bwd: y = d_hq_next (input gradient of visual_attn.linear_in of the NEXT step) -> exactly
 *              vln_envdrop_state_bwd(d_hc, d_xh_next, d_hq_next = y, htilde, apply_tanh = 1, d_src, ...).
 * `counters`: device uint32 [16], zero before the first launch (the kernel returns them to zero); launches that share
 * a counter buffer must be ordered on one stream. */
int vln_linear_state_fwd(const void* w_hi, const void* w_lo, int N, int K, const float* x, int ldx, int M, float* y,
                         int ldy, float* xh_next, int ld_xh, float* hq_next, float* hc_cur, float p,
                         const uint64_t* rng, uint64_t off_q, uint64_t off_c, unsigned int* counters, void* stream);
int vln_linear_state_bwd(const void* w_hi, const void* w_lo, int N, int K, const float* x, int ldx, int M, float* y,
                         int ldy, const float* d_hc, const float* d_xh_next, int ld_dxh, const float* htilde,
                         int ld_h, float* d_src, float p, const uint64_t* rng, uint64_t off_q, uint64_t off_c,
                         unsigned int* counters, void* stream);

/* Glue of the fused EnvDrop decoder step (EnvDropDecoder.forward policy.py:208-246): everything between
 * two grid-wide kernels of the step, writing into the operand rows of the next GEMM.
 * state_fwd: h~ = apply_tanh ? tanh(src) : src  (src = linear_out pre-activation, units.py:120, or the
 *   encoder's decoder_init);  xh_next[b, 0:H) (rows ld_xh apart) = h~ (LSTM hidden input of the NEXT step,
 *   policy.py:238);  hq_next [B,H] = dropout(h~; p, off_q) (next step's prev_h1_drop, policy.py:233);
 *   hc_cur [B,H] = dropout(h~; p, off_c) (this step's h_tilde_drop, policy.py:243).  Outputs nullable.
 * state_bwd: d_src = (dropout_c'(d_hc) + d_xh_next + dropout_q'(d_hq_next)) * (apply_tanh ? 1-h~^2 : 1);
 *   each gradient input nullable; htilde rows ld_h apart.
 * act_fwd: act [B,E] = tanh(W_a angle128(pose4[view]) + b_a) (policy.py:222, envdrop.py:76-78) — `w` is
 *   W_a reduced to its four group sums [E,4] (w[j][k] = sum_i W_a[j][32k+i]: the angle feature repeats each of
 *   its 4 values 32x, misc.py:286-293) — and
 *   xh[b, 0:E) = dropout(act; p, call_off).   act_bwd: d_actpre = dropout'(d_xh[:, 0:E)) * (1 - act^2). */
int vln_envdrop_state_fwd(const float* src, int apply_tanh, float* xh_next, int ld_xh, float* hq_next,
                          float* hc_cur, int B, int H, float p, const uint64_t* rng, uint64_t off_q,
                          uint64_t off_c, void* stream);
int vln_envdrop_state_bwd(const float* d_hc, const float* d_xh_next, int ld_dxh, const float* d_hq_next,
                          const float* htilde, int ld_h, int apply_tanh, float* d_src, int B, int H, float p,
                          const uint64_t* rng, uint64_t off_q, uint64_t off_c, void* stream);
int vln_envdrop_act_fwd(const int32_t* view, const float* pose4, const float* w, const float* bias, float* act,
                        float* xh, int ld_xh, int B, int E, float p, const uint64_t* rng, uint64_t call_off,
                        void* stream);
/* act_bwd covers n_steps steps in one launch: d_xh [n_steps,B,ld_dxh], act / d_actpre [n_steps,B,E]; step t
 * regenerates its mask from stream call_off + t*off_stride. */
int vln_envdrop_act_bwd(const float* d_xh, int ld_dxh, const float* act, float* d_actpre, int B, int E,
                        int n_steps, float p, const uint64_t* rng, uint64_t call_off, uint64_t off_stride,
                        void* stream);
/* vln_policy_fwd + vln_env_step + vln_envdrop_act_fwd (for the NEW view) in one launch, one warp per
 * episode: the tail of a rollout step (envdrop.py:177-219) and the head of the next (policy.py:222-223).
 * xh may be NULL (no following decoder pass): then the action embedding is skipped.
 * feedback = mode (0 teacher, 1 argmax, 2 sample) | (teacher_from + 1) << 8: with the upper field set, episodes
 * b >= teacher_from follow the teacher whatever the mode — the teacher-forced and the sampled rollout of one
 * EnvDrop iteration (trainer.py:411-421) stepped as one batch. */
int vln_policy_env_act_fwd(const float* logits, const int32_t* target, int feedback, const uint64_t* rng,
                           uint64_t off_sample, float* ce, int32_t* action, float* logp, float* entropy,
                           float* probs, const int32_t* vp_in, const int32_t* view_in, const uint8_t* ended_in,
                           const float* dist_in, const int32_t* goal, const int32_t* cand_vp,
                           const int32_t* cand_view, const int32_t* n_cand, const int32_t* next_hop,
                           const float* dist_tbl, const int64_t* sq_off, const int32_t* vp_local,
                           int32_t* vp_out, int32_t* view_out, uint8_t* ended_out, float* dist_out,
                           int32_t* teacher_out, float* reward, float* mask, int32_t* n_active,
                           const float* pose4, const float* w_act, const float* b_act, float* act, float* xh,
                           int ld_xh, int E, float p_act, uint64_t off_act, int B, void* stream);
/* vln_cand_logits_fwd (without bias) + vln_policy_env_act_fwd in ONE launch, one CTA per episode: the whole tail of
 * a fused decoder step (policy.py:199-206, envdrop.py:166-219, policy.py:222-223).  Bit-identical results; the
 * transition's index loads are issued ahead of / in parallel with the candidate rows instead of as a dependent chain.
 * `logits` [B,16] is written as well (saved for the backward pass). */
int vln_cand_policy_env_act_fwd(const vln_ctx* ctx, const int32_t* vp_in, const int32_t* view_in, const float* cand_ang4,
                                const float* tgt, float* logits, float drop_p, uint64_t off_cand, const int32_t* target,
                                int feedback, const uint64_t* rng, uint64_t off_sample, float* ce, int32_t* action,
                                float* logp, float* entropy, float* probs, const uint8_t* ended_in, const float* dist_in,
                                const int32_t* goal, const int32_t* cand_vp, const int32_t* cand_view,
                                const int32_t* n_cand, const int32_t* next_hop, const float* dist_tbl,
                                const int64_t* sq_off, const int32_t* vp_local, int32_t* vp_out, int32_t* view_out,
                                uint8_t* ended_out, float* dist_out, int32_t* teacher_out, float* reward, float* mask,
                                int32_t* n_active, const float* pose4, const float* w_act, const float* b_act, float* act,
                                float* xh, int ld_xh, int E, float p_act, uint64_t off_act, int B, void* stream);
/* vln_policy_bwd folded into vln_cand_logits_bwd: dlogits are computed from the action head's saved
 * probs / target / action / entropy and the incoming g_ce / g_logp / g_ent (each nullable).  Covers
 * n_steps decoder steps in one launch (none of its inputs depends on the backward recursion): every array
 * is [n_steps, B, ...]; step s regenerates its candidate mask from stream call_off + s*off_stride. */
int vln_cand_logits_bwd_policy(const vln_ctx* ctx, const int32_t* vp, const int32_t* view,
                               const int32_t* cand_view, const float* cand_ang4, const int32_t* n_cand,
                               const float* probs, const int32_t* target, const int32_t* action,
                               const float* entropy, const float* g_ce, const float* g_logp,
                               const float* g_ent, float* d_tgt, int B, int n_steps, float drop_p,
                               const uint64_t* rng, uint64_t call_off, uint64_t off_stride, void* stream);

/* Skinny linear layer on tcgen05 tensor cores (nn.Linear / nn.LSTMCell gate GEMMs of policy.py and
 * units.py at batch sizes <= 128):  y[m,n] (+)= sum_k x[m,k] w[n,k] (+ bias[n]),  m < M <= 128.
 * w_hi / w_lo: the weight [N,K] split into bf16 hi + lo (vln_split_bf16); x fp32 [M,K] row stride ldx;
 * y fp32 row stride ldy: overwritten (accumulate = 0) or added to (accumulate = 1, e.g. x W_ih^T + h W_hh^T).
 * K % 64 == 0, N % 4 == 0.  splits <= 0 picks the split-K factor (a power of two <= 8: the splits of one
 * 128-row weight tile are one thread-block cluster and reduce through distributed shared memory — no
 * atomics, deterministic).  bf16x3 (hi.hi + hi.lo + lo.hi, fp32 accumulate in TMEM) ~ fp32 accuracy.
 * For input gradients pass the transposed split (w^T as a [K,N] weight) and x = dY. */
int vln_linear_bf16x3(const void* w_hi, const void* w_lo, int N, int K, const float* x, int ldx, int M,
                      const float* bias, float* y, int ldy, int accumulate, int splits, void* stream);
/* The same product for TALL activations (M > 128 rows): the encoder's input projection over all B*L token
 * rows (nn.LSTM's x W_ih^T + b, units.py:58-63) and the critic over all T*B states (policy.py:263).
 * One CTA per (128 weight rows x 128 activation rows) block over the whole K, plain stores (or += with
 * accumulate = 1); same bf16x3 precision, same operand rules. */
int vln_linear_bf16x3_tall(const void* w_hi, const void* w_lo, int N, int K, const float* x, int ldx, int M,
                           const float* bias, float* y, int ldy, int accumulate, void* stream);
/* Two independent products of identical shape in ONE launch (y0 += x0 W0^T, y1 += x1 W1^T; both
 * accumulate): the candidate projection of step t (policy.py:199-206) and the visual-attention query
 * of step t+1 (units.py:107) both only wait for h~_t. */
int vln_linear_bf16x3_pair(const void* w0_hi, const void* w0_lo, const float* x0, float* y0, const void* w1_hi,
                           const void* w1_lo, const float* x1, float* y1, int N, int K, int ldx, int M, int ldy,
                           void* stream);
/* fp32 w [N,K] -> bf16 hi, lo [N,K] and, if hi_t/lo_t are given, the transposed pair [K,N]. */
int vln_split_bf16(const float* w, void* hi, void* lo, void* hi_t, void* lo_t, int N, int K, void* stream);

/* Persistent length-masked (Bi)LSTM recurrence: the packed-sequence nn.LSTM of EncoderLSTM
 * (units.py:58-71) for one layer, both directions in one launch.  Host arrays of n_dir (1 or 2)
 * device pointers: xproj[k] [B,L,4H] = x W_ih^T + b_ih + b_hh, w_hh[k] [4H,H] (gate order i,f,g,o),
 * acts[k] [B,L,4H] and cs[k] [B,L,H] (saved activations / cell states for backward).  Direction k
 * writes columns [kH,(k+1)H) of out [B,L,n_dir*H] (pre-zeroed; stays zero past each row's length)
 * and of h_last / c_last [B,n_dir*H] (state at each row's last valid step); direction 1 runs
 * reversed in time.  H per direction must be 128 or 256.  A cluster of H/32 CTAs owns 8 batch rows:
 * W_hh slices stay in shared memory across timesteps, h_t is exchanged through DSMEM. */
int vln_lstm_seq_fwd(const float* const* xproj, const float* const* w_hh, const int32_t* lengths, float* out,
                     float* const* acts, float* const* cs, float* h_last, float* c_last, int B, int L, int H,
                     int n_dir, void* stream);
/* Backward through time: d_out [B,L,n_dir*H], d_hlast / d_clast [B,n_dir*H] (each nullable) ->
 * d_xproj[k] [B,L,4H] (pre-zeroed; masked steps stay zero).  dW_hh, dW_ih, dx are GEMMs on d_xproj. */
int vln_lstm_seq_bwd(const float* const* w_hh, const int32_t* lengths, const float* const* acts,
                     const float* const* cs, const float* d_out, const float* d_hlast, const float* d_clast,
                     float* const* d_xproj, int B, int L, int H, int n_dir, void* stream);
/* Which recurrence kernels the two entry points above run: 1 = tcgen05 (W_hh resident in tensor memory, 16 or 32
 * batch rows per cluster; csrc/lstm_tc.cu, the default), 0 = mma.sync (weights in registers, 8 rows per cluster;
 * csrc/lstm_seq.cu), -1 = back to the VLN_LSTM_VARIANT environment default.  A tuning / testing knob. */
int vln_lstm_set_variant(int variant);

/* Action head (envdrop.py:166-195, follower.py:107-135, monitor.py:143-176): masked
 * log-softmax, CE(ignore_index=-1), argmax / Philox-sampled / teacher action, log-prob and
 * entropy of the chosen action.  logits [B,16] with -inf beyond the valid slots.
 *   feedback: 0 teacher, 1 argmax, 2 sample.  Outputs per episode: ce, action (int32),
 *   logp (of action), entropy, probs[B,16] (saved for backward). */
int vln_policy_fwd(const float* logits, const int32_t* target, int feedback, const uint64_t* rng,
                   uint64_t call_off, float* ce, int32_t* action, float* logp, float* entropy,
                   float* probs, int B, void* stream);
/* dlogits[b,j] = g_ce[b]*(p - onehot(target)) + g_logp[b]*(onehot(action) - p)
 *               - g_ent[b]*p*(log p + H)   (each g_* nullable) */
int vln_policy_bwd(const float* probs, const int32_t* target, const int32_t* action,
                   const float* entropy, const float* g_ce, const float* g_logp, const float* g_ent,
                   float* dlogits, int B, void* stream);

/* nn.Dropout on a dense fp32 tensor with the library's Philox stream (fwd and bwd are the same
 * op: y = x * keep / (1-p)). */
int vln_dropout(const float* x, float* y, int64_t n, float p, const uint64_t* rng, uint64_t call_off,
                void* stream);
/* keep-mask bytes (1 = kept) for elements [0,n) of the stream (seed, offset). */
int vln_dropout_mask(uint8_t* mask, int64_t n, float p, const uint64_t* rng, uint64_t call_off, void* stream);
/* nn.Embedding(padding_idx) + nn.Dropout of EncoderLSTM.forward (units.py:48-52) in one pass over the
 * [rows = B*L, E] token rows: y = drop(emb[tokens]) with the keep mask vln_dropout_mask gives for the dense
 * [rows, E] tensor under (rng, call_off); p = 0 is a plain lookup.  The backward writes the whole d_emb [V,E]
 * (rows of one vocabulary entry summed in row order: deterministic; the padding row is zero). */
int vln_embed_drop_fwd(const int64_t* tokens, const float* emb, float* y, int64_t rows, int E, int V, float p,
                       const uint64_t* rng, uint64_t call_off, void* stream);
int vln_embed_drop_bwd(const int64_t* tokens, const float* d_y, float* d_emb, int64_t rows, int E, int V,
                       int padding_idx, float p, const uint64_t* rng, uint64_t call_off, void* stream);
/* rng[1] += delta (one thread); lets graph replays move to fresh streams. */
int vln_rng_advance(uint64_t* rng, uint64_t delta, void* stream);

/* Stub-simulator transition + observation on index tables (EnvBatch.makeActions
 * common_env.py:91-110, R2RBatch.observe :299-330, _teacher_action base.py:159-178, reward
 * shaping envdrop.py:207-219).  Out of place: state row t -> state row t+1, so a rollout keeps
 * its whole [T+1,B] trajectory on the device (backward kernels re-read vp/view of each step).
 *   action[b] in 0..n_cand (n_cand = STOP) or -1.  Outputs for the NEW state: teacher[b]
 *   (-1 if ended), dist[b]; for the transition: reward[b] (EnvDrop shaping), mask[b] = !ended
 *   before the step.  n_active (nullable): atomically += number of episodes still running
 *   after the step (the device-side `ended.all()` of envdrop.py:219). */
int vln_env_step(const int32_t* vp_in, const int32_t* view_in, const uint8_t* ended_in,
                 const float* dist_in, const int32_t* goal, const int32_t* action,
                 const int32_t* cand_vp, const int32_t* cand_view, const int32_t* n_cand,
                 const int32_t* next_hop, const float* dist_tbl, const int64_t* sq_off,
                 const int32_t* vp_local, int32_t* vp_out, int32_t* view_out, uint8_t* ended_out,
                 float* dist_out, int32_t* teacher, float* reward, float* mask, int32_t* n_active,
                 int B, void* stream);
/* teacher/dist for the current state without moving (reset). */
int vln_env_observe(const int32_t* vp, const uint8_t* ended, const int32_t* goal,
                    const int32_t* cand_vp, const int32_t* n_cand, const int32_t* next_hop,
                    const float* dist_tbl, const int64_t* sq_off, const int32_t* vp_local,
                    int32_t* teacher, float* dist, int B, void* stream);

/* Feature of the action just taken (follower.py:164, monitor.py:191: a_t_prev = cands[i, max(a,0)]):
 * the agent first rewrites STOP / ignored / already-ended actions to -1 (follower.py:141-146), so
 * slot = (ended[b] || action[b] < 0 || action[b] >= n_cand) ? 0 : action[b]; out[b] = candidate
 * row `slot` of state (vp[b], view[b]), all-zero if the viewpoint has no candidate.  `ended`
 * (state BEFORE the step) may be NULL.  fp32 [B,2176], bit-exact. */
int vln_gather_action_feat(const vln_ctx* ctx, const int32_t* vp, const int32_t* view,
                           const int32_t* action, const uint8_t* ended, const int32_t* cand_view, const float* cand_ang4,
                           const int32_t* n_cand, float* out, int B, void* stream);

/* A2C loss assembly (envdrop.py:240-264).  Inputs are time-major [T,B]: reward, mask (fp32),
 * logp, entropy, value (critic(hidden_states[t])), plus last_value[B] (critic(last_h), no grad)
 * and ended[B] after the last step.  Discounted returns run backwards in time in float64, the
 * first multiply by gamma in float32 (numpy promotion at envdrop.py:238-243).  Outputs:
 *   loss_b[B] = sum_t mask*( -logp*(R-v) + 0.5*(R-v)^2 - ent_coef*entropy )   (advantage detached)
 *   ret[T,B]  = the returns R_t (saved for backward), total[1] += sum mask, critic_sq[1] += sum mask*(R-v)^2 */
int vln_a2c_fwd(const float* reward, const float* mask, const float* logp, const float* entropy,
                const float* value, const float* last_value, const uint8_t* ended, float gamma,
                float ent_coef, float* loss_b, float* ret, float* total, float* critic_sq, int T, int B,
                void* stream);
/* g_b[B] = d(total loss)/d(loss_b).  d_logp = -g*mask*(R-v); d_value = -g*mask*(R-v); d_entropy = -g*mask*ent_coef. */
int vln_a2c_bwd(const float* g_b, const float* mask, const float* value, const float* ret,
                float ent_coef, float* d_logp, float* d_value, float* d_entropy, int T, int B,
                void* stream);

/* Gradient clip + optimiser in one pass over flat fp32 buffers (trainer.py:423-427:
 * clip_grad_norm_(encoder,40), clip_grad_norm_(decoder,40), critic unclipped, RMSprop/Adam).
 * The flat buffer is laid out as n_groups (<=4) contiguous groups, group k = [group_off[k],
 * group_off[k+1]) (host array of n_groups+1 offsets); max_norm[k] <= 0 means "not clipped"
 * (host array).  Offsets must be multiples of 4 floats and the buffers 16-byte aligned (16-byte accesses).
 * sqnorm [n_groups] is device scratch holding each group's squared L2 norm of
 * grad*grad_scale.  kind 0 = RMSprop(alpha=.99, eps=1e-8), 1 = Adam(.9,.999,1e-8), torch semantics;
 * step is the 1-based Adam step count. */
int vln_grad_sqnorm(const float* grad, const int64_t* group_off, int n_groups, float* sqnorm,
                    float grad_scale, void* stream);
int vln_optim_step(float* param, const float* grad, float* state1, float* state2,
                   const int64_t* group_off, const float* max_norm, int n_groups,
                   const float* sqnorm, float grad_scale, int kind, float lr, int step, void* stream);

/* Weight gradient of every nn.Linear / nn.LSTMCell of the path (what autograd computes as dY.t() @ X for policy.py:238,
 * units.py:58-60, 107, 119): dw[M,N] (+)= sum_r dy[r,m] * x[r,n] over the R stacked rows of a rollout, on tcgen05
 * (kind::tf32, both operands MN-major straight from TMA, fp32 accumulation in tensor memory).  dy [R, ld_dy], x [R, ld_x],
 * dw [M, ld_dw] fp32, M / N / strides multiples of 4.  accumulate != 0 adds into dw (e.g. a .grad buffer).  `scratch`
 * (scratch_floats fp32) is optional: with it small outputs are split over row ranges so that the launch fills the SMs; the
 * partial blocks are merged in range order by a second launch (the result does not depend on scheduling). */
int vln_wgrad_tf32(const float* dy, int ld_dy, const float* x, int ld_x, int R, int M, int N, float* dw, int ld_dw,
                   int accumulate, float* scratch, int64_t scratch_floats, void* stream);

/* Chain regions.  The kernels of a decoder step (policy.py:208-246 and its backward) form a chain of grid-wide
 * dependencies; between vln_chain_begin and vln_chain_end consecutive launches on `stream` resolve them through counters in
 * `flags` (n_flags uint32 in device memory, zeroed here by a memset on the stream; the last one counts poll time-outs and must
 * read 0) instead of waiting for the predecessor grid to complete (same results).  Regions do not
 * nest; a launch on another stream, or of a kernel that is not link-aware, simply falls back to the plain dependency.
 * vln_chain_end returns the number of counters used.  An experiment kept behind VLN_CHAIN_FLAGS=1 (default off: on B200
 * griddepcontrol.wait measured faster, 4.42 vs 4.81 ms per EnvDrop iteration); with it off both calls do nothing. */
int vln_chain_begin(unsigned int* flags, int n_flags, void* stream);
int vln_chain_end(void);

/* Input gradient of a tall nn.Linear (autograd's dY @ W for units.py:58-60, the encoder's input projection, whose dx only
 * feeds the embedding gradient): dx[M,N] = dy[M,R] w[R,N], same tcgen05 kind::tf32 kernel with dy as the K-major operand.
 * fp32, R / N / strides multiples of 4. */
int vln_dgrad_tf32(const float* dy, int ld_dy, const float* w, int ld_w, int M, int R, int N, float* dx, int ld_dx,
                   void* stream);

/* Gradient of the instruction context accumulated over the n decoder steps of a rollout (autograd through
 * SoftDotAttention, units.py:107-118, n times): out[b,l,j] (+)= sum_t a[t,b,l] * v[t,b,j], fp32 FMAs.
 * a at a + t*a_step + b*lda + l (l < L), v at v + t*v_step + b*ldv + j (j < H), out [B,L,H] contiguous. */
int vln_seq_outer_sum(const float* a, int64_t a_step, int lda, const float* v, int64_t v_step, int ldv, int n, int B,
                      int L, int H, float* out, int accumulate, void* stream);

/* Evaluation.score (src/engine/evaluator.py:41-146; DTW src/utils/dtw.py:60-82; CLS src/utils/cls.py:62-90) for N
 * trajectories in one launch, float64 on the fp32 all-pairs distance table of the environment kernels.
 * pred int32 [N,P] / ref int32 [N,R] hold global viewpoint indices (first pred_len[n] / ref_len[n] entries valid, R <= 16);
 * out float64 [N,8] = {nav_error, oracle_error, steps, trajectory length, SPL term, nDTW, SDTW, CLS}. */
int vln_eval_paths(const int32_t* pred, const int32_t* pred_len, int P, const int32_t* ref, const int32_t* ref_len,
                   int R, const float* dist_tbl, const int64_t* sq_off, const int32_t* vp_local, double margin,
                   double* out, int N, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VLN_B200_H_ */
