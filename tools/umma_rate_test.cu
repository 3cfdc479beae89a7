// Experiment: tensor-pipe cost (clk) of one tcgen05.mma kind::f16 M=128 x N x K=16, SS mode, as a function of N,
// the smem layout (SWIZZLE_128B / 64B / 32B rows), window alignment (aligned vs row-shifted start, SBO = 8*rb vs a
// patch pitch) and whether consecutive MMAs re-read the same rows (k-steps inside a 128-byte row) or new rows.
// One CTA per SM on all SMs (smem bandwidth is per SM); one elected thread issues R MMAs, commits, waits.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../domain-transfer-gan_b200/csrc/common.cuh"
using namespace dtg;

struct Cfg { int N, rb, layout, a_sbo, a_shift_rows, ksteps, R, M; };

template <int KS, int PITCH, int RB, int SHIFT>
__global__ void __launch_bounds__(128, 1) k(Cfg c, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~uintptr_t(1023));
  uint64_t* bar = (uint64_t*)(smem + 160 * 1024);
  uint32_t* slot = (uint32_t*)(bar + 2);
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  int warp = threadIdx.x >> 5;
  if (warp == 0) {
    if ((threadIdx.x & 31) == 0) { mbar_init(&bar[0], 1); mbar_fence_init(); }
    __syncwarp();
    tmem_alloc(slot, 512);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  uint32_t tm = *slot;
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    const uint32_t idesc = umma_idesc(1, 0, 0, c.M, c.N);
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 96 * 1024);
    const uint32_t a_hi = ((uint32_t)c.a_sbo >> 4) | (1u << 14) | ((uint32_t)c.layout << 29);
    const uint32_t b_hi = ((8u * c.rb) >> 4) | (1u << 14) | ((uint32_t)c.layout << 29);
    const uint32_t a_lo0 = ((sa >> 4) & 0x3FFF) | (1u << 16), b_lo0 = ((sb >> 4) & 0x3FFF) | (1u << 16);
    t0 = clock64();
    if (elect_one()) {
      // compile-time offsets: every descriptor is base + constant, independent of the previous MMA
      tc_mma<false>(tm, ((uint64_t)a_hi << 32) | a_lo0, ((uint64_t)b_hi << 32) | b_lo0, idesc, 0);
      const uint32_t nb16 = (uint32_t)(c.N * RB) >> 4;
      for (int rep = 0; rep < c.R / (9 * KS); ++rep) {
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const uint32_t a_lo = a_lo0 + ((((t / 3) * PITCH + t % 3 + SHIFT) * RB) >> 4);
          const uint32_t b_lo = b_lo0 + t * nb16;
#pragma unroll
          for (int j = 0; j < KS; ++j)
            tc_mma<false>(tm, ((uint64_t)a_hi << 32) | (a_lo + 2 * j), ((uint64_t)b_hi << 32) | (b_lo + 2 * j), idesc, 1);
        }
      }
      tc_commit(&bar[0]);
    }
    __syncwarp();
    mbar_wait(&bar[0], 0);
    t1 = clock64();
    tc_fence_after();
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  const int R = 2304;
  int Ns[] = {16, 32, 64, 128, 256};
#define RUN(NAME, KS, PITCH, RB, LAYOUT, SHIFT, MM)                                                              \
  for (int N : Ns) {                                                                                         \
    if (9 * N * RB > 64 * 1024) continue;                                                                    \
    Cfg c{N, RB, LAYOUT, PITCH * RB, SHIFT, KS, R, MM};                                                      \
    k<KS, PITCH, RB, SHIFT><<<148, 128, 180 * 1024>>>(c, d);                                                 \
    cudaError_t e = cudaDeviceSynchronize();                                                                 \
    if (e != cudaSuccess) { printf("%s N=%d: CUDA error %s\n", NAME, N, cudaGetErrorString(e)); return 1; }  \
    long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);                                   \
    long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;                              \
    printf("%s M=%3d N=%3d: %.1f clk/MMA (floor %d)\n", NAME, MM, N, (double)mx / (R / (9 * KS) * 9 * KS), N / 2 > 8 ? N / 2 : 8); \
  }
  cudaFuncSetAttribute(k<4, 8, 128, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k<1, 8, 128, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k<4, 10, 128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k<1, 10, 128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k<2, 8, 64, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k<2, 10, 64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k<1, 8, 32, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k<1, 14, 32, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  RUN("sw128 aligned  ks4", 4, 8, 128, 2, 0, 64)
  RUN("sw128 aligned  ks4", 4, 8, 128, 2, 0, 128)
  RUN("sw128 aligned  ks1", 1, 8, 128, 2, 0, 128)
  RUN("sw128 patch10  ks4", 4, 10, 128, 2, 1, 128)
  RUN("sw128 patch10  ks1", 1, 10, 128, 2, 1, 128)
  RUN("sw64  aligned  ks2", 2, 8, 64, 4, 0, 128)
  RUN("sw64  patch10  ks2", 2, 10, 64, 4, 1, 128)
  RUN("sw32  aligned  ks1", 1, 8, 32, 6, 0, 128)
  RUN("sw32  patch14  ks1", 1, 14, 32, 6, 3, 128)
  return 0;
}
