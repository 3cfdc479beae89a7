// Shared device/host helpers for the dtg_b200 sm_100a kernels: PTX wrappers for mbarrier, TMA,
// tcgen05 (MMA / TMEM / commit), UMMA descriptor builders and error plumbing.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/dtg_b200.h"

namespace dtg {

// ------------------------------------------------------------------------------------------
// error plumbing (thread-local last error; entry points never throw / abort)
// ------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
int cu_fail(CUresult e, const char* what);

#define DTG_CHECK_CUDA(expr)                                  \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) return ::dtg::cuda_fail(_e, #expr); \
  } while (0)

// every kernel launch site: bump the library-wide launch counter, then surface launch errors
void count_launch();
#define DTG_LAUNCHED()                         \
  do {                                         \
    ::dtg::count_launch();                     \
    DTG_CHECK_CUDA(cudaGetLastError());        \
  } while (0)

// Kernel launch with programmatic dependent launch (PDL): the next kernel in the stream may begin launching and run
// its prologue while this one drains; every kernel calls pdl_enter() (trigger + wait) before touching global memory,
// so data dependencies (RAW and WAR) are still honoured.  DTG_NO_PDL=1 disables the attribute.
bool pdl_enabled();
int norm_impl();               // dtg_set_option("norm_impl"): 0 cluster-fused, 1 TMA-staged, 2 two-phase streaming
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
  if (e == cudaSuccess) count_launch();
  return e;
}
#endif

#define DTG_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      ::dtg::set_error(__VA_ARGS__);      \
      return DTG_ERR_INVALID;             \
    }                                     \
  } while (0)

// Tensor-map encode through the driver entry point (no link-time libcuda dependency).
int encode_tiled(CUtensorMap* map, int dtype, int rank, void* base, const uint64_t* dims,
                 const uint64_t* strides_bytes /* rank-1 entries */, const uint32_t* box,
                 int swizzle /* 0 none, 1 = 128B (16-byte atoms), 2 = 128B with 32-byte atoms, 3 = 64B, 4 = 32B */);

static inline int elem_size(int dtype) { return dtype == DTG_BF16 ? 2 : 4; }

// Shared-memory budget (bytes) of ONE tensor-core CTA.  The tcgen05 kernels are persistent, one CTA per SM; what they
// leave free decides whether a bandwidth-bound kernel of another stream (normalisation, losses, optimizer: ~22 KB of
// static shared memory per CTA) can be co-resident on the same SM and stream from HBM while the tensor pipe works.
// dtg_set_option("smem_cap_kb") / DTG_SMEM_CAP_KB.
int tensor_smem_budget();

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__

// PDL: let the dependent grid start launching, then wait until every prerequisite grid has completed and its memory
// is visible.  Must precede the first global-memory access of a kernel.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() {
  pdl_trigger();
  pdl_wait();
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---- TMA ----
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "r"(c2), "r"(c3)
      : "memory");
}

// ---- tcgen05 / TMEM ----
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// arrives on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
template <bool TF32>
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                       uint32_t accumulate) {
  if constexpr (TF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread (thread = lane)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- UMMA descriptors (SWIZZLE_128B, 128-byte rows, 1024-byte 8-row atoms) ----
// K-major operand (rows = M/N index, 128 B of K per row): SBO = 1024 between 8-row groups.
// MN-major operand (rows = K index, 128 B of M/N per row): LBO = byte stride between 128-B-wide
// M/N blocks, SBO = 1024 between 8-row (K) groups.
// layout_type: 2 = SWIZZLE_128B (16-byte swizzle atoms); 1 = SWIZZLE_128B_BASE32B (32-byte atoms, the
// only layout tcgen05 accepts for MN-major 32-bit (tf32) operands: 4-row K groups, SBO = 512).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                    uint32_t layout_type = 2) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;  // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(layout_type) << 61;
  return d;
}
// instruction descriptor: fp32 accumulate, A/B format (1 = bf16, 2 = tf32), majors (0 = K, 1 = MN)
__host__ __device__ __forceinline__ uint32_t umma_idesc(uint32_t fmt, uint32_t a_mn, uint32_t b_mn, uint32_t M,
                                                        uint32_t N) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// fp32 planes feed tcgen05 kind::tf32, which TRUNCATES the low 13 mantissa bits of its operands.
// Producers therefore store round-to-nearest tf32 values, which removes the truncation bias.
__device__ __forceinline__ float round_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case DTG_ACT_RELU: return v > 0.f ? v : 0.f;
    case DTG_ACT_LRELU: return v > 0.f ? v : 0.2f * v;
    case DTG_ACT_TANH: return tanhf(v);
    default: return v;
  }
}

// Mirror destinations for a reflect halo: interior index i in [0,L) also lands at -i (1<=i<=halo)
// and at 2(L-1)-i (L-1-halo <= i <= L-2).  Returns count; fills idx[] (interior first).
__device__ __forceinline__ int reflect_targets(int i, int L, int halo, int (&idx)[3]) {
  int n = 0;
  idx[n++] = i;
  if (halo > 0) {
    if (i >= 1 && i <= halo) idx[n++] = -i;
    if (i <= L - 2 && i >= L - 1 - halo) idx[n++] = 2 * (L - 1) - i;
  }
  return n;
}

#endif  // __CUDACC__
}  // namespace dtg
