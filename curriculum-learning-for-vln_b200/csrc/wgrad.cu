// Weight gradients on the 5th-generation tensor cores:   dW[M, N] (+)= dY^T X  =  sum_r dY[r, m] * X[r, n]
//
// over the R = T*B stacked rows of a whole rollout (the autograd side of every nn.Linear / nn.LSTMCell of the path:
// policy.py:238, units.py:58-60, 107, 119; what torch.autograd computes as dY.t() @ X).  Both operands are row-major
// with the REDUCTION dimension as their slow axis, i.e. "MN-major" for the MMA: tcgen05.mma.kind::tf32 reads them
// straight from swizzled shared-memory tiles (128-byte span, 32-byte chunks: the one layout tcgen05 takes for MN-major
// 32-bit operands; TMA's SWIZZLE_128B_ATOM_32B writes it) that TMA fills from the fp32 buffers — no transposition, no
// conversion pass — with fp32 accumulation in tensor memory.  TF32 inputs (10-bit mantissa) over 2.7-10 k-long sums:
// gradient cosine vs the fp32 oracle stays >= 0.9999 (tests/test_kernels_gpu.py::test_wgrad_tcgen05, rollout tests).
//
// CTA = one 128 x 128 block of dW (grid.x = M tiles, grid.y = N tiles), optionally one of `splits` row ranges
// (grid.z) when the block count alone would leave most SMs idle; warp 0 = TMA producer, warp 1 = MMA issuer,
// warps 4-7 = epilogue (tcgen05.ld -> global).  3-stage ring of 64-row chunks: per stage 4 + 4 TMA boxes of
// [64 rows x 32 floats] (one 128-byte swizzle row per reduction row), 8 MMAs of 128 x 128 x 8.
// Rows past R are zero-filled by TMA (they add nothing); columns past M / N are computed on zeros and never stored.
// Merge of the row ranges is DETERMINISTIC: every range writes its partial block to a scratch slab and a second small
// launch adds the slabs in range order into dW (all loads in flight at once; a last-arriver merge inside the GEMM ran the
// 16 dependent slab reads of a block on one CTA and cost more than the product).  Without splits the block goes to dW directly.
#include <cstdio>
#include <cstdlib>
#include <cudaTypedefs.h>

#include "common.cuh"

int vln_make_tmap_2d_f32_sw(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                            uint32_t box_cols, uint32_t box_rows, int atom32);

namespace {

constexpr int kBM = 128, kBN = 128, kBK = 64;           // block of dW; reduction rows per stage
constexpr int kStages = 3;
constexpr int kThreads = 256;
constexpr int kChunkBytes = kBK * 128;                  // one [64 rows x 32 floats] box
constexpr int kTileBytes = 4 * kChunkBytes;             // 128 columns = 4 boxes = 32 KB per operand
constexpr int kStageBytes = 2 * kTileBytes;             // A + B
constexpr int kSmem = kStages * kStageBytes + 1024 + 256;

// MN-major, SWIZZLE_128B_BASE32B: 32-float (128-byte) groups along M/N are LBO apart, 4-row groups (one 512-byte swizzle
// atom) along K are SBO apart
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr) {
  constexpr uint32_t lbo = kChunkBytes, sbo = 512;
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;           // leading-dimension byte offset: next 32-float group
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;           // stride byte offset: next group of 8 reduction rows
  d |= (uint64_t)1 << 46;                               // descriptor version (sm_100)
  d |= (uint64_t)1 << 61;                               // SWIZZLE_128B_BASE32B
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit_(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tmem_ld32_(uint32_t taddr, uint32_t (&v)[32]) {
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

// K-major, SWIZZLE_128B: rows of 32 floats (128 bytes), 8-row groups SBO = 1024 bytes apart (operand A of the NN product)
__device__ __forceinline__ uint64_t make_desc_k(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                               // SWIZZLE_128B
  return d;
}

// A_K = false: C[M,N] = A^T B with A [R,M], B [R,N] (weight gradient; both MN-major).
// A_K = true : C[M,N] = A B    with A [M,R] (K-major: the reduction is its fast axis), B [R,N] (input gradient dY W).
template <bool A_K>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_tf32_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, float* __restrict__ c,
                  int ldc, int M, int N, int R, int splits, float* __restrict__ scratch, int accumulate) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + kStages * kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* acc_done = bars + 2 * kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * kBN, split = blockIdx.z;
  const int nkb = (R + kBK - 1) / kBK;
  const int kb0 = (int)((long long)split * nkb / splits), kb1 = (int)((long long)(split + 1) * nkb / splits);
  const int n_iter = kb1 - kb0;

  if (tid == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_done, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kBN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (n_iter > 0) {
    if (warp == 0) {
      if (lane == 0) {
        for (int it = 0; it < n_iter; ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          uint8_t* st = base + (size_t)s * kStageBytes;
          mbar_expect_tx(&full[s], kStageBytes);
          const int r0 = (kb0 + it) * kBK;
          if (A_K) {                                           // 2 boxes of [128 rows x 32 reduction columns]
            tma_load_2d(st, &tm_a, &full[s], r0, m0);
            tma_load_2d(st + kTileBytes / 2, &tm_a, &full[s], r0 + 32, m0);
          }
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            if (!A_K) tma_load_2d(st + ch * kChunkBytes, &tm_a, &full[s], m0 + 32 * ch, r0);
            tma_load_2d(st + kTileBytes + ch * kChunkBytes, &tm_b, &full[s], n0 + 32 * ch, r0);
          }
        }
      }
    } else if (warp == 1) {
      // D = F32, A = B = TF32, both MN-major, N = 128, M = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_K ? 0u : 1u) << 15) | (1u << 16) |
                             ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
      const uint32_t base_addr = smem_u32(base);
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1u;
        mbar_wait(&full[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_addr = base_addr + (uint32_t)s * kStageBytes, b_addr = a_addr + kTileBytes;
        const uint64_t da = A_K ? make_desc_k(a_addr) : make_desc_mn(a_addr), db = make_desc_mn(b_addr);
#pragma unroll
        for (int k = 0; k < kBK / 8; ++k) {
          const uint64_t ko = (uint64_t)(k * (1024 >> 4));           // 8 reduction rows of 128 bytes
          // K-major A: 8 reduction columns = 32 bytes inside the swizzled 128-byte row; the second box after 4 steps
          const uint64_t ka = A_K ? (uint64_t)(((k >> 2) * (kTileBytes / 2) + (k & 3) * 32) >> 4) : ko;
          umma_tf32(tmem_d, da + ka, db + ko, idesc, (it | k) != 0);
        }
        umma_commit_(&empty[s]);
      }
      umma_commit_(acc_done);
      __syncwarp();
    } else if (warp >= 4) {
      mbar_wait(acc_done, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int wq = warp & 3;
      const int m = m0 + wq * 32 + lane;                      // TMEM lane = row of the dW block
      float* dst_row = (splits > 1 ? scratch + ((size_t)split * M + m) * (size_t)N : c + (size_t)m * ldc);
#pragma unroll 1
      for (int cb = 0; cb < kBN / 32; ++cb) {
        uint32_t v[32];
        tmem_ld32_(tmem_d + ((uint32_t)(wq * 32) << 16) + (uint32_t)(cb * 32), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (m < M) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const int n = n0 + cb * 32 + j;
            if (n >= N) break;                                 // (N is a multiple of 4)
            float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                   __uint_as_float(v[j + 3]));
            float4* p = reinterpret_cast<float4*>(dst_row + n);
            if (splits == 1 && accumulate) {
              const float4 old = *p;
              o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
            }
            *p = o;
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(kBN) : "memory");
  }
}

// Deterministic merge of the row ranges: dw (+)= slab_0 + slab_1 + ... in range order, one float4 per thread, the loads of
// all ranges in flight together.
template <int kMax>
__global__ void __launch_bounds__(256)
wgrad_merge_kernel(const float* __restrict__ scratch, float* __restrict__ c, int ldc, int M, int N, int splits, int accumulate) {
  pdl_wait();
  const int n4 = N >> 2;
  const long long item = (long long)blockIdx.x * 256 + threadIdx.x;
  if (item >= (long long)M * n4) return;
  const int m = (int)(item / n4), n = (int)(item % n4) * 4;
  const float* src = scratch + (size_t)m * N + n;
  float4 part[kMax];
#pragma unroll
  for (int s = 0; s < kMax; ++s)
    if (s < splits) part[s] = __ldcg(reinterpret_cast<const float4*>(src + (size_t)s * M * N));
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int s = 0; s < kMax; ++s)
    if (s < splits) {
      acc.x += part[s].x; acc.y += part[s].y; acc.z += part[s].z; acc.w += part[s].w;
    }
  float4* dst = reinterpret_cast<float4*>(c + (size_t)m * ldc + n);
  if (accumulate) {
    const float4 old = *dst;
    acc.x += old.x; acc.y += old.y; acc.z += old.z; acc.w += old.w;
  }
  *dst = acc;
}

}  // namespace

extern "C" int vln_wgrad_tf32(const float* dy, int ld_dy, const float* x, int ld_x, int R, int M, int N, float* dw, int ld_dw,
                              int accumulate, float* scratch, int64_t scratch_floats, void* stream) {
  VLN_REQUIRE(dy && x && dw && R > 0 && M > 0 && N > 0, "bad arguments");
  VLN_REQUIRE(M % 4 == 0 && N % 4 == 0 && ld_dy % 4 == 0 && ld_x % 4 == 0 && ld_dw % 4 == 0, "M, N and row strides must be multiples of 4");
  VLN_REQUIRE((((uintptr_t)dy | (uintptr_t)x | (uintptr_t)dw) & 15) == 0, "operands must be 16-byte aligned");
  VLN_REQUIRE(ld_dy >= M && ld_x >= N && ld_dw >= N, "row strides too small");
  CUtensorMap tm_a, tm_b;
  int rc = vln_make_tmap_2d_f32_sw(&tm_a, dy, (uint64_t)R, (uint64_t)M, (uint64_t)ld_dy, 32, kBK, 1);
  if (rc) return rc;
  rc = vln_make_tmap_2d_f32_sw(&tm_b, x, (uint64_t)R, (uint64_t)N, (uint64_t)ld_x, 32, kBK, 1);
  if (rc) return rc;
  const int tiles_m = (M + kBM - 1) / kBM, tiles_n = (N + kBN - 1) / kBN;
  const int nkb = (R + kBK - 1) / kBK;
  // row ranges: only when the blocks alone leave most of the 148 SMs idle and there is a scratch slab to merge through
  int splits = 1;
  if (scratch) {
    const int tiles = tiles_m * tiles_n;
    splits = tiles <= 74 ? 148 / tiles : (tiles < 148 ? 296 / tiles : 1);   // one or two full waves of CTAs
    if (splits > nkb / 4) splits = nkb / 4;                      // at least 4 chunks per range
    if (splits > 16) splits = 16;
    if (splits < 1) splits = 1;
    if ((int64_t)splits * M * N > scratch_floats) splits = 1;
  }
  static bool configured = false;
  if (!configured) {
    VLN_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tf32_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    configured = true;
  }
  wgrad_tf32_kernel<false><<<dim3(tiles_m, tiles_n, splits), kThreads, kSmem, (cudaStream_t)stream>>>(
      tm_a, tm_b, dw, ld_dw, M, N, R, splits, scratch, accumulate);
  VLN_LAUNCH_OK();
  if (splits > 1) {
    const long long items = (long long)M * (N >> 2);
    VLN_CHECK_CUDA(vln_launch_chain(wgrad_merge_kernel<16>, dim3((unsigned)((items + 255) / 256)), dim3(256), 0, (cudaStream_t)stream,
                                    (const float*)scratch, dw, ld_dw, M, N, splits, accumulate));
  }
  return 0;
}

// dX[M,N] = dY[M,R] W[R,N]: the input gradient of a tall nn.Linear (the encoder's input projection: its dx only feeds the
// embedding gradient), same kernel with operand A K-major.
extern "C" int vln_dgrad_tf32(const float* dy, int ld_dy, const float* w, int ld_w, int M, int R, int N, float* dx, int ld_dx,
                              void* stream) {
  VLN_REQUIRE(dy && w && dx && R > 0 && M > 0 && N > 0, "bad arguments");
  VLN_REQUIRE(R % 4 == 0 && N % 4 == 0 && ld_dy % 4 == 0 && ld_w % 4 == 0 && ld_dx % 4 == 0, "R, N and row strides must be multiples of 4");
  VLN_REQUIRE((((uintptr_t)dy | (uintptr_t)w | (uintptr_t)dx) & 15) == 0, "operands must be 16-byte aligned");
  VLN_REQUIRE(ld_dy >= R && ld_w >= N && ld_dx >= N, "row strides too small");
  CUtensorMap tm_a, tm_b;
  int rc = vln_make_tmap_2d_f32_sw(&tm_a, dy, (uint64_t)M, (uint64_t)R, (uint64_t)ld_dy, 32, kBM, 0);
  if (rc) return rc;
  rc = vln_make_tmap_2d_f32_sw(&tm_b, w, (uint64_t)R, (uint64_t)N, (uint64_t)ld_w, 32, kBK, 1);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    VLN_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tf32_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    configured = true;
  }
  wgrad_tf32_kernel<true><<<dim3((M + kBM - 1) / kBM, (N + kBN - 1) / kBN, 1), kThreads, kSmem, (cudaStream_t)stream>>>(
      tm_a, tm_b, dx, ld_dx, M, N, R, 1, nullptr, 0);
  VLN_LAUNCH_OK();
  return 0;
}

// ---- d_ctx: out[b, l, j] (+)= sum_t a[t, b, l] * v[t, b, j] -----------------------------------------------------------
// The gradient of the instruction context through the n decoder steps of a rollout (units.py:107-118 applied n times:
// d_ctx[b] = sum_t attn_t[b]^T (x) d_weighted_t[b]; and dCW[b] = sum_t dlogit_t[b]^T (x) drop(h_1)_t[b]): n is 35-70, far too
// short a reduction for the tensor cores, so plain fp32 FMAs: CTA = (128 columns j, 40 rows l, one episode b), the [n, 40]
// coefficient slab in shared memory (broadcast reads), one column per thread with 40 accumulators.
namespace {
constexpr int kOL = 40, kOT = 8;
template <int C>                                               // columns per thread (2 when H and the strides are even)
__global__ void __launch_bounds__(128)
seq_outer_kernel(const float* __restrict__ a, long long a_st, int lda, const float* __restrict__ v, long long v_st, int ldv,
                 int n, int L, int H, float* __restrict__ out, int accumulate) {
  extern __shared__ float s_a[];                               // [n rounded up to kOT][kOL], zero rows past n
  const int b = blockIdx.x, j = (blockIdx.y * 128 + threadIdx.x) * C, l0 = blockIdx.z * kOL;
  const int n_pad = (n + kOT - 1) / kOT * kOT;
  for (int i = threadIdx.x; i < n_pad * kOL; i += 128) {
    const int t = i / kOL, l = l0 + i % kOL;
    s_a[i] = (t < n && l < L) ? a[(size_t)t * a_st + (size_t)b * lda + l] : 0.f;
  }
  __syncthreads();
  if (j >= H) return;
  float acc[kOL][C];
#pragma unroll
  for (int l = 0; l < kOL; ++l)
#pragma unroll
    for (int c = 0; c < C; ++c) acc[l][c] = 0.f;
  const float* vp = v + (size_t)b * ldv + j;
  for (int t0 = 0; t0 < n; t0 += kOT) {
    float x[kOT][C];
#pragma unroll
    for (int u = 0; u < kOT; ++u) {
      if (t0 + u < n) {
        if (C == 2) {
          const float2 p = __ldg(reinterpret_cast<const float2*>(vp + (size_t)(t0 + u) * v_st));
          x[u][0] = p.x;
          x[u][C - 1] = p.y;
        } else {
          x[u][0] = __ldg(vp + (size_t)(t0 + u) * v_st);
        }
      } else {
#pragma unroll
        for (int c = 0; c < C; ++c) x[u][c] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < kOT; ++u) {
      const float4* row = reinterpret_cast<const float4*>(s_a + (t0 + u) * kOL);
#pragma unroll
      for (int q = 0; q < kOL / 4; ++q) {
        const float4 w = row[q];
#pragma unroll
        for (int c = 0; c < C; ++c) {
          acc[4 * q][c] = fmaf(w.x, x[u][c], acc[4 * q][c]);
          acc[4 * q + 1][c] = fmaf(w.y, x[u][c], acc[4 * q + 1][c]);
          acc[4 * q + 2][c] = fmaf(w.z, x[u][c], acc[4 * q + 2][c]);
          acc[4 * q + 3][c] = fmaf(w.w, x[u][c], acc[4 * q + 3][c]);
        }
      }
    }
  }
  float* o = out + ((size_t)b * L + l0) * H + j;
#pragma unroll
  for (int l = 0; l < kOL; ++l) {
    if (l0 + l < L) {
      if (C == 2) {
        float2* q = reinterpret_cast<float2*>(o + (size_t)l * H);
        float2 r = make_float2(acc[l][0], acc[l][C - 1]);
        if (accumulate) {
          const float2 old = *q;
          r.x += old.x;
          r.y += old.y;
        }
        *q = r;
      } else {
        o[(size_t)l * H] = accumulate ? o[(size_t)l * H] + acc[l][0] : acc[l][0];
      }
    }
  }
}
}  // namespace

extern "C" int vln_seq_outer_sum(const float* a, int64_t a_step, int lda, const float* v, int64_t v_step, int ldv, int n, int B,
                                 int L, int H, float* out, int accumulate, void* stream) {
  VLN_REQUIRE(a && v && out && n > 0 && B > 0 && L > 0 && H > 0, "bad arguments");
  VLN_REQUIRE(n <= 296, "at most 296 steps");
  const size_t smem = (size_t)((n + kOT - 1) / kOT * kOT) * kOL * sizeof(float);
  const bool two = H % 2 == 0 && ldv % 2 == 0 && v_step % 2 == 0 && (((uintptr_t)v | (uintptr_t)out) & 7) == 0;
  if (two)
    seq_outer_kernel<2><<<dim3(B, (H + 255) / 256, (L + kOL - 1) / kOL), 128, smem, (cudaStream_t)stream>>>(
        a, (long long)a_step, lda, v, (long long)v_step, ldv, n, L, H, out, accumulate);
  else
    seq_outer_kernel<1><<<dim3(B, (H + 127) / 128, (L + kOL - 1) / kOL), 128, smem, (cudaStream_t)stream>>>(
        a, (long long)a_step, lda, v, (long long)v_step, ldv, n, L, H, out, accumulate);
  VLN_LAUNCH_OK();
  return 0;
}
