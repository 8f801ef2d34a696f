// Soft-dot attention over the instruction context [B, L, H] (fp32), forward and backward.
// Same cluster scheme as the panorama kernel: S CTAs per episode each own H/S columns of the
// context tile (staged once in shared memory), per-row partial dots are summed through DSMEM.
// Roofline class: HBM/L2 — algorithmic bytes per episode-step: L*H*4 read forward;
// backward L*H*4 read + 2*L*H*4 read-modify-write of d_context.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxL = 96;

// mode 0: forward. mode 1: backward.
__global__ void __launch_bounds__(kThreads) ctx_attn_kernel(
    const float* __restrict__ context, const float* __restrict__ tgt, const int32_t* __restrict__ lengths,
    float* __restrict__ attn, float* __restrict__ weighted,            // fwd outputs
    const float* __restrict__ d_weighted, const float* __restrict__ d_attn_ext, float* __restrict__ d_tgt,
    float* __restrict__ d_context,                                      // bwd
    int mode, int L, int H, int HS, int ld_w) {
  extern __shared__ __align__(16) uint8_t smem[];
  float* tile = reinterpret_cast<float*>(smem);          // [L][HS]
  float* vsl = tile + (size_t)L * HS;                    // [HS]  tgt slice (fwd) / d_weighted slice (bwd)
  float* tsl = vsl + HS;                                 // [HS]  tgt slice (bwd only)
  float* part = tsl + HS;                                // [kMaxL]
  float* sm = part + kMaxL;                              // [kMaxL]
  float* av = sm + kMaxL;                                // [kMaxL] saved attn (bwd)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rank = (int)cluster_ctarank(), S = (int)cluster_nctarank();
  const int b = blockIdx.y;
  const int len = min(lengths[b], L);
  const int col0 = rank * HS;
  const int hv = HS / 4;                                 // float4 per row slice

  const float* cb = context + (size_t)b * L * H;
  for (int i = tid; i < len * hv; i += kThreads) {
    const int l = i / hv, c = i - l * hv;
    reinterpret_cast<float4*>(tile)[l * hv + c] = __ldg(reinterpret_cast<const float4*>(cb + (size_t)l * H + col0) + c);
  }
  const float* v_in = (mode == 0 ? tgt + (size_t)b * H : d_weighted + (size_t)b * ld_w) + col0;
  for (int i = tid; i < HS; i += kThreads) {
    vsl[i] = v_in[i];
    if (mode == 1) tsl[i] = tgt[(size_t)b * H + col0 + i];
  }
  __syncthreads();

  for (int l = warp; l < len; l += kThreads / 32) {
    float acc = 0.f;
    for (int c = lane; c < HS; c += 32) acc += tile[l * HS + c] * vsl[c];
    acc = warp_sum(acc);
    if (lane == 0) part[l] = acc;
  }
  cluster_arrive();
  cluster_wait();
  if (tid < len) {
    float tot = 0.f;
    for (int r = 0; r < S; ++r) tot += dsmem_ld_f32(part + tid, (uint32_t)r);
    sm[tid] = tot;
  }
  cluster_arrive();
  __syncthreads();

  if (warp == 0) {
    float x[3], a[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int l = lane + 32 * k;
      x[k] = (l < len) ? sm[l] : -INFINITY;
    }
    if (mode == 0) {
      const float m = warp_max(fmaxf(x[0], fmaxf(x[1], x[2])));
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        a[k] = (lane + 32 * k < len) ? expf(x[k] - m) : 0.f;
        s += a[k];
      }
      const float inv = 1.0f / warp_sum(s);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int l = lane + 32 * k;
        if (l < L) {
          const float p = (l < len) ? a[k] * inv : 0.f;
          if (l < len) sm[l] = p;
          if (rank == 0) attn[(size_t)b * L + l] = p;
        }
      }
    } else {
      float dot = 0.f, p[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int l = lane + 32 * k;
        p[k] = (l < len) ? attn[(size_t)b * L + l] : 0.f;
        if (l < len && d_attn_ext) x[k] += d_attn_ext[(size_t)b * L + l];
        dot += (l < len) ? p[k] * x[k] : 0.f;
      }
      dot = warp_sum(dot);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int l = lane + 32 * k;
        if (l < len) {
          sm[l] = p[k] * (x[k] - dot);   // dlogit
          av[l] = p[k];
        }
      }
    }
  }
  __syncthreads();

  if (mode == 0) {
    for (int c = tid; c < HS; c += kThreads) {
      float acc = 0.f;
      for (int l = 0; l < len; ++l) acc += sm[l] * tile[l * HS + c];
      weighted[(size_t)b * ld_w + col0 + c] = acc;
    }
  } else {
    for (int c = tid; c < HS; c += kThreads) {
      float acc = 0.f;
      for (int l = 0; l < len; ++l) acc += sm[l] * tile[l * HS + c];
      d_tgt[(size_t)b * H + col0 + c] = acc;
    }
    float* db = d_context + (size_t)b * L * H;
    for (int i = tid; d_context != nullptr && i < len * hv; i += kThreads) {
      const int l = i / hv, c = i - l * hv;
      float4* p = reinterpret_cast<float4*>(db + (size_t)l * H + col0) + c;
      float4 cur = *p;
      const float4 dw = reinterpret_cast<const float4*>(vsl)[c], tg = reinterpret_cast<const float4*>(tsl)[c];
      const float a = av[l], dlg = sm[l];
      cur.x += a * dw.x + dlg * tg.x;
      cur.y += a * dw.y + dlg * tg.y;
      cur.z += a * dw.z + dlg * tg.z;
      cur.w += a * dw.w + dlg * tg.w;
      *p = cur;
    }
  }
  cluster_wait();
}

int launch(const float* context, const float* tgt, const int32_t* lengths, float* attn, float* weighted,
           const float* d_weighted, const float* d_attn_ext, float* d_tgt, float* d_context, int mode, int B, int L,
           int H, int ld_w, cudaStream_t stream) {
  VLN_REQUIRE(ld_w >= H, "row stride of weighted / d_weighted must be >= H");
  VLN_REQUIRE(L > 0 && L <= kMaxL, "L must be in 1..96");
  VLN_REQUIRE(H % 16 == 0 && H >= 64, "H must be a multiple of 16");
  int S = 4;
  while (S > 1 && (H % (S * 4) != 0)) S >>= 1;
  const int HS = H / S;
  const size_t smem = ((size_t)L * HS + 2 * HS + 3 * kMaxL) * sizeof(float);
  static size_t configured = 0;
  if (smem > configured) {
    VLN_CHECK_CUDA(cudaFuncSetAttribute(ctx_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(S, B);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = S;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  VLN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, ctx_attn_kernel, context, tgt, lengths, attn, weighted, d_weighted,
                                    d_attn_ext, d_tgt, d_context, mode, L, H, HS, ld_w));
  return 0;
}

}  // namespace

extern "C" int vln_ctx_attn_fwd_ld(const float* context, const float* tgt, const int32_t* lengths, float* attn,
                                   float* weighted, int ld_weighted, int B, int L, int H, void* stream) {
  VLN_REQUIRE(context && tgt && lengths && attn && weighted && B > 0, "bad arguments");
  return launch(context, tgt, lengths, attn, weighted, nullptr, nullptr, nullptr, nullptr, 0, B, L, H, ld_weighted,
                (cudaStream_t)stream);
}

extern "C" int vln_ctx_attn_fwd(const float* context, const float* tgt, const int32_t* lengths, float* attn,
                                float* weighted, int B, int L, int H, void* stream) {
  return vln_ctx_attn_fwd_ld(context, tgt, lengths, attn, weighted, H, B, L, H, stream);
}

extern "C" int vln_ctx_attn_bwd_ld(const float* context, const float* tgt, const int32_t* lengths, const float* attn,
                                   const float* d_weighted, int ld_d_weighted, const float* d_attn_ext, float* d_tgt,
                                   float* d_context, int B, int L, int H, void* stream) {
  VLN_REQUIRE(context && tgt && lengths && attn && d_weighted && d_tgt && B > 0, "bad arguments");
  return launch(context, tgt, lengths, const_cast<float*>(attn), nullptr, d_weighted, d_attn_ext, d_tgt, d_context,
                1, B, L, H, ld_d_weighted, (cudaStream_t)stream);
}

extern "C" int vln_ctx_attn_bwd(const float* context, const float* tgt, const int32_t* lengths, const float* attn,
                                const float* d_weighted, const float* d_attn_ext, float* d_tgt, float* d_context,
                                int B, int L, int H, void* stream) {
  return vln_ctx_attn_bwd_ld(context, tgt, lengths, attn, d_weighted, H, d_attn_ext, d_tgt, d_context, B, L, H, stream);
}
