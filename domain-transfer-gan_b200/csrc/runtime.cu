// Error plumbing, version, and the cuTensorMapEncodeTiled trampoline (resolved through
// cudaGetDriverEntryPoint so the library has no link-time dependency on libcuda).
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>

#include "common.cuh"

namespace dtg {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- runtime options: -1 = not yet initialised from the environment -----------------------------------------------
struct Option {
  const char* key;
  const char* env;
  int env_value;      // value when the environment variable is set (integer variables: parsed instead)
  int dflt;
  bool env_is_int;
  std::atomic<int> v{-1};
};
static Option g_opts[] = {
    {"pdl", "DTG_NO_PDL", 0, 1, false},
    {"norm_impl", "DTG_NORM_IMPL", 0, 2, true},
    {"smem_cap_kb", "DTG_SMEM_CAP_KB", 0, 227, true},
};

static int opt_get(int i) {
  Option& o = g_opts[i];
  int v = o.v.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv(o.env);
    v = e ? (o.env_is_int ? atoi(e) : o.env_value) : o.dflt;
    if (v < 0) v = o.dflt;
    o.v.store(v, std::memory_order_relaxed);
  }
  return v;
}

bool pdl_enabled() { return opt_get(0) != 0; }
int norm_impl() { return opt_get(1); }

int tensor_smem_budget() {
  int kb = opt_get(2);
  if (kb < 96) kb = 96;
  if (kb > 227) kb = 227;
  return kb * 1024;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", static_cast<int>(e), cudaGetErrorString(e), what);
  return DTG_ERR_CUDA;
}

int cu_fail(CUresult e, const char* what) {
  set_error("CUDA driver error %d at %s", static_cast<int>(e), what);
  return DTG_ERR_CUDA;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int encode_tiled(CUtensorMap* map, int dtype, int rank, void* base, const uint64_t* dims,
                 const uint64_t* strides_bytes, const uint32_t* box, int swizzle) {
  EncodeTiledFn fn = get_encode();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    return DTG_ERR_CUDA;
  }
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gs[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = fn(map, dtype == DTG_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                  static_cast<cuuint32_t>(rank), base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle == 1   ? CU_TENSOR_MAP_SWIZZLE_128B
                  : swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                  : swizzle == 3 ? CU_TENSOR_MAP_SWIZZLE_64B
                  : swizzle == 4 ? CU_TENSOR_MAP_SWIZZLE_32B
                                 : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u] base %p",
              static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
              rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, base);
    return DTG_ERR_CUDA;
  }
  return DTG_OK;
}

}  // namespace dtg

extern "C" int dtg_version(void) { return DTG_VERSION; }

extern "C" int dtg_set_option(const char* key, int value) {
  if (key != nullptr && value >= 0)
    for (size_t i = 0; i < sizeof(dtg::g_opts) / sizeof(dtg::g_opts[0]); ++i)
      if (strcmp(key, dtg::g_opts[i].key) == 0) {
        const int prev = dtg::opt_get(static_cast<int>(i));
        dtg::g_opts[i].v.store(value, std::memory_order_relaxed);
        return prev;
      }
  dtg::set_error("dtg_set_option: unknown key or negative value");
  return DTG_ERR_INVALID;
}

extern "C" unsigned long long dtg_launch_count(void) { return dtg::g_launches.load(std::memory_order_relaxed); }

extern "C" int dtg_last_error(char* buf, size_t cap) {
  size_t n = strlen(dtg::g_err);
  if (buf && cap > 0) {
    size_t k = n < cap - 1 ? n : cap - 1;
    memcpy(buf, dtg::g_err, k);
    buf[k] = 0;
  }
  return static_cast<int>(n);
}
