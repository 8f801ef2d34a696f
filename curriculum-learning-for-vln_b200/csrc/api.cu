// C-ABI plumbing: errors, context (feature table + TMA descriptor).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cudaTypedefs.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void vln_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool vln_pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("VLN_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

// ---- chain links (common.cuh): host-side bookkeeping of one region at a time ----
namespace {
struct ChainState {
  bool active = false, prev_valid = false;
  cudaStream_t stream = nullptr;
  unsigned int* flags = nullptr;
  int n = 0, next = 0;
  unsigned int prev_signals = 0;
} g_chain;
bool chain_flags_enabled() {
  static int on = -1;
  if (on < 0) {
    // off unless VLN_CHAIN_FLAGS=1: measured on B200 the counters lose to griddepcontrol.wait (4.81 vs 4.42 ms per
    // iteration): the hardware releases a dependent grid ~0.8 us after the primary's last CTA retires, a fence + atomic
    // + polling round trip costs more than that; most of the "gap" between two step kernels is the primary's own
    // reduction traffic draining
    const char* e = getenv("VLN_CHAIN_FLAGS");
    on = (e && e[0] == '1') ? 1 : 0;
  }
  return on != 0 && vln_pdl_enabled();
}
}  // namespace

ChainLink vln_chain_link(cudaStream_t stream, unsigned int signals) {
  ChainLink l{nullptr, nullptr, nullptr, 0u};
  if (!g_chain.active || stream != g_chain.stream) return l;
  if (g_chain.next >= g_chain.n - 1) {                       // out of counters: plain dependencies from here on
    g_chain.prev_valid = false;
    return l;
  }
  l.err = g_chain.flags + (g_chain.n - 1);
  if (g_chain.prev_valid) {
    l.wait_flag = g_chain.flags + (g_chain.next - 1);
    l.wait_n = g_chain.prev_signals;
  }
  l.done_flag = g_chain.flags + g_chain.next;
  g_chain.next++;
  g_chain.prev_signals = signals;
  g_chain.prev_valid = true;
  return l;
}
void vln_chain_break(cudaStream_t stream) {
  if (g_chain.active && stream == g_chain.stream) g_chain.prev_valid = false;
}

extern "C" int vln_chain_begin(unsigned int* flags, int n_flags, void* stream) {
  VLN_REQUIRE(flags && n_flags >= 16, "need an array of at least 16 counters");
  g_chain.active = false;
  if (!chain_flags_enabled()) return 0;
  VLN_CHECK_CUDA(cudaMemsetAsync(flags, 0, (size_t)n_flags * sizeof(unsigned int), (cudaStream_t)stream));
  g_chain.active = true;
  g_chain.prev_valid = false;
  g_chain.stream = (cudaStream_t)stream;
  g_chain.flags = flags;
  g_chain.n = n_flags;
  g_chain.next = 0;
  return 0;
}
extern "C" int vln_chain_end(void) {
  const int used = g_chain.active ? g_chain.next : 0;
  g_chain.active = false;
  g_chain.prev_valid = false;
  return used;
}

extern "C" const char* vln_last_error(void) { return g_err; }
extern "C" int vln_version(void) { return 100; }

static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
  return fn;
}

// Generic 2-D bf16 row-major tensor map; used by the feature table and by the GEMM operands.
int vln_make_tmap_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                     uint32_t box_cols, uint32_t box_rows, int swizzle128) {
  auto fn = get_encode_fn();
  if (!fn) {
    vln_set_error("cuTensorMapEncodeTiled entry point not available");
    return -4;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = CUDA_SUCCESS;
  for (int attempt = 0; attempt < 2; ++attempt) {
    r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
           CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    // a thread that has made no runtime call yet (an autograd worker's first kernel) has no current driver context:
    // bind the primary context through the runtime and encode again
    if (r != CUDA_ERROR_INVALID_CONTEXT) break;
    cudaFree(0);
  }
  if (r != CUDA_SUCCESS) {
    vln_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu ld=%llu box=%ux%u)", (int)r,
                  (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld_elems, box_cols,
                  box_rows);
    return -5;
  }
  return 0;
}

// 2-D fp32 row-major tensor map without swizzle (raw activation tiles of the skinny GEMM).
int vln_make_tmap_2d_f32(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                         uint32_t box_cols, uint32_t box_rows) {
  auto fn = get_encode_fn();
  if (!fn) {
    vln_set_error("cuTensorMapEncodeTiled entry point not available");
    return -4;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {ld_elems * 4};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = CUDA_SUCCESS;
  for (int attempt = 0; attempt < 2; ++attempt) {
    r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstr, box, estr,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_ERROR_INVALID_CONTEXT) break;
    cudaFree(0);                                               // (see vln_make_tmap_2d)
  }
  if (r != CUDA_SUCCESS) {
    vln_set_error("cuTensorMapEncodeTiled(f32) failed with CUresult %d (rows=%llu cols=%llu ld=%llu box=%ux%u)", (int)r,
                  (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld_elems, box_cols, box_rows);
    return -5;
  }
  return 0;
}

// 2-D fp32 row-major tensor map with a 128-byte swizzle (box_cols = 32 floats = one swizzle row): MN-major tf32 operands
// of the weight-gradient kernel (csrc/wgrad.cu).  atom32: 32-byte chunks are permuted within the 128-byte span (the only
// shared-memory layout tcgen05 takes for MN-major 32-bit operands), else the common 16-byte chunks.
int vln_make_tmap_2d_f32_sw(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                            uint32_t box_cols, uint32_t box_rows, int atom32) {
  auto fn = get_encode_fn();
  if (!fn) {
    vln_set_error("cuTensorMapEncodeTiled entry point not available");
    return -4;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {ld_elems * 4};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = CUDA_SUCCESS;
  for (int attempt = 0; attempt < 2; ++attempt) {
    r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstr, box, estr,
           CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_ERROR_INVALID_CONTEXT) break;
    cudaFree(0);                                               // (see vln_make_tmap_2d)
  }
  if (r != CUDA_SUCCESS) {
    vln_set_error("cuTensorMapEncodeTiled(f32, swizzle 128B) failed with CUresult %d (rows=%llu cols=%llu ld=%llu box=%ux%u)",
                  (int)r, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld_elems, box_cols, box_rows);
    return -5;
  }
  return 0;
}

extern "C" int vln_ctx_create(vln_ctx** out, const void* table_bf16, int n_vp, int device) {
  VLN_REQUIRE(out && table_bf16 && n_vp > 0, "null table or n_vp <= 0");
  VLN_REQUIRE(((uintptr_t)table_bf16 & 15) == 0, "table must be 16-byte aligned");
  VLN_CHECK_CUDA(cudaSetDevice(device));
  vln_ctx* c = new vln_ctx();
  c->table = (const __nv_bfloat16*)table_bf16;
  c->n_vp = n_vp;
  c->device = device;
  cudaDeviceProp prop;
  VLN_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  c->num_sms = prop.multiProcessorCount;
  int rc = vln_make_tmap_2d(&c->tmap_tile, table_bf16, (uint64_t)n_vp * VLN_V, VLN_IMG, VLN_IMG, 256, VLN_V, 0);
  if (rc) {
    delete c;
    return rc;
  }
  c->scratch = nullptr;
  c->tickets = nullptr;
  if (cudaMalloc(&c->scratch, (size_t)VLN_SPLIT_MAX_B * 4 * 2056 * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&c->tickets, (size_t)VLN_SPLIT_MAX_B * sizeof(unsigned int)) != cudaSuccess ||
      cudaMemset(c->tickets, 0, (size_t)VLN_SPLIT_MAX_B * sizeof(unsigned int)) != cudaSuccess) {
    vln_set_error("vln_ctx_create: cannot allocate the split-unit scratch: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(c->scratch);
    cudaFree(c->tickets);
    delete c;
    return -2;
  }
  *out = c;
  return 0;
}

extern "C" void vln_ctx_destroy(vln_ctx* ctx) {
  if (!ctx) return;
  cudaFree(ctx->scratch);
  cudaFree(ctx->tickets);
  delete ctx;
}
