// Shared device helpers for the sm_100a kernels: Philox keep-masks, mbarrier/TMA/cluster PTX,
// warp reductions, error plumbing for the C-ABI.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vln_b200.h"

// ---- error plumbing -----------------------------------------------------------------------
void vln_set_error(const char* fmt, ...);
#define VLN_CHECK_CUDA(expr)                                                          \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      vln_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return -2;                                                                      \
    }                                                                                 \
  } while (0)
#define VLN_REQUIRE(cond, msg)                                                        \
  do {                                                                                \
    if (!(cond)) {                                                                    \
      vln_set_error("%s: requirement failed: %s", __func__, msg);                     \
      return -1;                                                                      \
    }                                                                                 \
  } while (0)
#define VLN_LAUNCH_OK()                                                               \
  do {                                                                                \
    cudaError_t _e = cudaGetLastError();                                              \
    if (_e != cudaSuccess) {                                                          \
      vln_set_error("%s: launch failed: %s", __func__, cudaGetErrorString(_e));       \
      return -3;                                                                      \
    }                                                                                 \
  } while (0)

// ---- programmatic dependent launch (PDL) ------------------------------------------------------
// The decoder step is a chain of ~20 short kernels, each a grid-wide dependency of the next.  Kernels of
// the chain are launched with programmaticStreamSerialization: a kernel may start as soon as every CTA of
// its predecessor has executed pdl_trigger() (first statement of each kernel), runs its prologue
// (barrier init, TMEM allocation, weight-tile TMA) concurrently with the predecessor's tail and blocks in
// pdl_wait() before it touches anything a predecessor writes.  VLN_PDL=0 turns the attribute off.
bool vln_pdl_enabled();

// ---- chain links: grid-to-grid dependencies resolved by device-side counters ---------------------------------------
// Between two dependent kernels of the decoder step, griddepcontrol.wait returns only after the predecessor grid has
// COMPLETED and its memory has been flushed: measured 2-3.5 us from the predecessor's last store to the successor's
// release (tools/chain_stamps.py), 14 of the 38 us of a decoder step.  Inside a chain region (vln_chain_begin ...
// vln_chain_end on one stream) consecutive launches are linked through a zeroed array of counters instead: every CTA
// (or warp) of the producer adds 1 to its launch's counter after its last global access (fence + atomic), and the
// successor — already resident, it was launched programmatically — polls that counter with ld.acquire.gpu until all
// `wait_n` arrivals are in.  A launch that is not link-aware breaks the chain (its successor falls back to
// griddepcontrol.wait, which is always correct).  The counters live in caller memory and are zeroed by a memset node at
// the head of the region, so a captured graph replays the protocol unchanged.
struct ChainLink {
  const unsigned int* wait_flag;   // predecessor's counter, nullptr: griddepcontrol.wait
  unsigned int* done_flag;         // this launch's counter, nullptr: nothing to signal
  unsigned int* err;               // incremented when a poll times out (a protocol bug; read by vln_chain_end)
  unsigned int wait_n;             // arrivals the predecessor's launch produces
};
// `signals`: how many arrivals THIS launch will add to its counter (CTAs, or warps, as the kernel signals)
ChainLink vln_chain_link(cudaStream_t stream, unsigned int signals);
void vln_chain_break(cudaStream_t stream);

template <typename K, typename... Args>
inline cudaError_t vln_launch_linked(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = vln_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}
// launch of a kernel that does not take part in the link protocol
template <typename K, typename... Args>
inline cudaError_t vln_launch_chain(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  vln_chain_break(stream);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = vln_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

struct vln_ctx {
  const __nv_bfloat16* table;
  int n_vp;
  int device;
  int num_sms;
  CUtensorMap tmap_tile;   // [n_vp*36, 2048] bf16, box {256 cols, 36 rows}: one panorama slice
  float* scratch;          // [VLN_SPLIT_MAX_B * 4 * 2056] partial results of split panorama units
  unsigned int* tickets;   // [VLN_SPLIT_MAX_B] arrival counters (return to 0 after every launch)
};

// streaming variant of the panorama attention (pano_stream.cu), selected by vln_pano_attn_ld
cudaError_t vln_pano_stream_launch(const vln_ctx* ctx, const int32_t* vp, const int32_t* view, const float* loc4,
                                   const float* vec, int ld_vec, float* attn_io, float* out, int ld_out, int B, int mode,
                                   float drop_p, const uint8_t* mask_bits, cudaStream_t stream);

// ---- Philox4x32-10 ---------------------------------------------------------------------------
struct Philox8 {
  uint32_t w[4];  // 8 x 16-bit lanes
};

__host__ __device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#ifdef __CUDA_ARCH__
  uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
#else
  uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
  uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
  uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
  uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

// 128 random bits for block `blk` (= element index / 8) of stream (seed, offset).
__host__ __device__ __forceinline__ Philox8 philox8(uint64_t seed, uint64_t offset, uint64_t blk) {
  uint32_t c[4] = {(uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  Philox8 o;
  o.w[0] = c[0]; o.w[1] = c[1]; o.w[2] = c[2]; o.w[3] = c[3];
  return o;
}

__host__ __device__ __forceinline__ uint32_t drop_threshold(float p) {
  float t = p * 65536.0f + 0.5f;
  return t <= 0.f ? 0u : (t >= 65536.f ? 65536u : (uint32_t)t);
}
// lane j (0..7) of a Philox8 block: kept iff its 16-bit value >= thr
__host__ __device__ __forceinline__ bool philox_keep(const Philox8& r, int j, uint32_t thr) {
  uint32_t v = (r.w[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu;
  return v >= thr;
}
// uniform in [0,1) from 24 bits of word j (0..3)
__host__ __device__ __forceinline__ float philox_uniform(const Philox8& r, int j) {
  return (float)(r.w[j] >> 8) * (1.0f / 16777216.0f);
}

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// ---- chain links, device side ----
__device__ __forceinline__ void chain_spin(const ChainLink& c) {
  const long long t0 = clock64();
  unsigned int v;
  while (true) {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c.wait_flag) : "memory");
    if (v >= c.wait_n) break;
    if (clock64() - t0 > (1ll << 31)) {                      // ~1 s: never in a correct chain; do not hang the device
      atomicAdd(c.err, 1u);
      break;
    }
  }
}
// replaces a pdl_wait() that every thread of the CTA executes at the same point
__device__ __forceinline__ void chain_wait_cta(const ChainLink& c) {
  if (c.wait_flag == nullptr) {
    pdl_wait();
    return;
  }
  if (threadIdx.x == 0) chain_spin(c);
  __syncthreads();
}
// replaces a pdl_wait() executed by ONE thread that then issues TMA / bulk copies of the predecessor's output
__device__ __forceinline__ void chain_wait_thread(const ChainLink& c) {
  if (c.wait_flag == nullptr) {
    pdl_wait();
    return;
  }
  chain_spin(c);
  asm volatile("fence.proxy.async;" ::: "memory");           // generic-proxy observation before async-proxy reads
}
// after the CTA's last global access (all threads, converged): one arrival per CTA
__device__ __forceinline__ void chain_signal_cta(const ChainLink& c) {
  if (c.done_flag == nullptr) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async;" ::: "memory");         // bulk reductions / stores of this CTA (async proxy)
    __threadfence();
    atomicAdd(c.done_flag, 1u);
  }
}
// one arrival per WARP (kernels whose warps retire at different points)
__device__ __forceinline__ void chain_signal_warp(const ChainLink& c) {
  if (c.done_flag == nullptr) return;
  __syncwarp();
  if ((threadIdx.x & 31) == 0) {
    __threadfence();
    atomicAdd(c.done_flag, 1u);
  }
}
// ---- chain stamps (debug builds only: -DVLN_CHAIN_STAMPS, tools/chain_stamps.py) ---------------------------------
// Block 0 / thread 0 of a step-chain kernel records {kernel id, globaltimer at its first instruction, after the
// programmatic-dependency wait, at its last instruction} behind the rng state (ops.Rng allocates the room: word 2 =
// enable flag, word 3 = running count, 4 words per record from word 4), so the residency / release / work phases of
// the chain can be read off one clock across kernels.  Compiled out of the product library.
#ifdef VLN_CHAIN_STAMPS
__device__ __forceinline__ unsigned long long chain_gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned long long* chain_begin(const uint64_t* rng, int kid) {
  if (rng == nullptr || blockIdx.x != 0 || blockIdx.y != 0 || blockIdx.z != 0 || threadIdx.x != 0) return nullptr;
  unsigned long long* r = (unsigned long long*)rng;
  if (r[2] == 0ull) return nullptr;
  unsigned long long* s = r + 4 + (atomicAdd(&r[3], 1ull) % 8192ull) * 4ull;
  s[0] = (unsigned long long)kid;
  s[1] = chain_gtimer();
  s[2] = 0ull;
  s[3] = 0ull;
  return s;
}
#define CHAIN_BEGIN(rng, kid) unsigned long long* chain_s_ = chain_begin(rng, kid)
#define CHAIN_MARK(i)                          \
  do {                                         \
    if (chain_s_) chain_s_[i] = chain_gtimer(); \
  } while (0)
#else
#define CHAIN_BEGIN(rng, kid) \
  do {                        \
  } while (0)
#define CHAIN_MARK(i) \
  do {                \
  } while (0)
#endif

// ---- warp helpers ------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- shared-memory addresses, mbarrier, TMA, cluster ----------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking probe (try_wait may suspend the thread for a hardware time slice)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 2-D tiled TMA load: box at (c0 = inner/column coordinate, c1 = row coordinate)
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// read a float from the same smem offset in CTA `rank` of the cluster
__device__ __forceinline__ float dsmem_ld_f32(const float* local_ptr, uint32_t rank) {
  uint32_t a = smem_u32(local_ptr), ra;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
  return v;
}
#endif  // __CUDACC__
