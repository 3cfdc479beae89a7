// Experiment: can a K-major SWIZZLE_128B UMMA A-descriptor start at an arbitrary 128-byte ROW offset inside a
// TMA-loaded buffer (i.e. not 1024-byte aligned), and which base_offset encoding makes it correct?
// A buffer: 256 rows x 64 bf16 (A[r][k]); B: 64 rows x 64 (identity-like pattern); D[m][n] = sum_k A[m+off][k] B[n][k].
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
#include "../domain-transfer-gan_b200/csrc/common.cuh"
using namespace dtg;

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                            float* out, int off, int mode) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                 // 256 rows * 128 B = 32 KB
  uint8_t* sB = smem + 32768;         // 64 rows * 128 B
  uint64_t* bar = (uint64_t*)(smem + 32768 + 8192);
  uint32_t* slot = (uint32_t*)(bar + 2);
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    if (lane == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
    __syncwarp();
    tmem_alloc(slot, 64);
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  uint32_t tm = *slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar[0], 32768 + 8192);
    tma_load_2d(sA, &tmA, &bar[0], 0, 0);
    tma_load_2d(sA + 16384, &tmA, &bar[0], 0, 128);
    tma_load_2d(sB, &tmB, &bar[0], 0, 0);
  }
  mbar_wait(&bar[0], 0);
  tc_fence_after();
  if (threadIdx.x == 0) {
    uint32_t idesc = umma_idesc(1, 0, 0, 128, 64);
    uint32_t a0 = smem_u32(sA) + off * 128;
    for (int j = 0; j < 4; ++j) {
      uint64_t ad = umma_desc_sw128(a0 + j * 32, 16, mode == 2 ? 1280 : 1024);
      if (mode == 1) ad |= (uint64_t)((a0 >> 7) & 7) << 49;     // base_offset = row phase within the 1024-B atom
      uint64_t bd = umma_desc_sw128(smem_u32(sB) + j * 32, 16, 1024);
      tc_mma<false>(tm, ad, bd, idesc, j > 0);
    }
    tc_commit(&bar[1]);
  }
  mbar_wait(&bar[1], 0);
  tc_fence_after();
  int row = warp * 32 + lane;
  for (int c0 = 0; c0 < 64; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[row * 64 + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 64); }
}

int main() {
  std::vector<__nv_bfloat16> hA(256 * 64), hB(64 * 64);
  for (int r = 0; r < 256; ++r) for (int c = 0; c < 64; ++c) hA[r * 64 + c] = __float2bfloat16((float)((r * 7 + c * 3) % 17) - 8.f);
  for (int n = 0; n < 64; ++n) for (int c = 0; c < 64; ++c) hB[n * 64 + c] = __float2bfloat16((float)((n * 5 + c) % 7) - 3.f);
  __nv_bfloat16 *dA, *dB; float* dOut;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dOut, 128 * 64 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tA, tB;
  uint64_t dimsA[2] = {64, 256}, strA[1] = {128}; uint32_t boxA[2] = {64, 128};
  uint64_t dimsB[2] = {64, 64}; uint32_t boxB[2] = {64, 64};
  if (encode_tiled(&tA, DTG_BF16, 2, dA, dimsA, strA, boxA, 1) || encode_tiled(&tB, DTG_BF16, 2, dB, dimsB, strA, boxB, 1)) { printf("encode failed\n"); return 1; }
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  std::vector<float> h(128 * 64);
  for (int mode = 0; mode < 3; ++mode)
    for (int off = 0; off <= 40; off += (off < 9 ? 1 : 11)) {
      k<<<1, 128, 48 * 1024, 0>>>(tA, tB, dOut, off, mode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d off %d: CUDA error %s\n", mode, off, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h.data(), dOut, h.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0;
      for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
        double ref = 0;
        const int arow = mode == 2 ? off + (m / 8) * 10 + (m % 8) : m + off;
        for (int c = 0; c < 64; ++c) ref += (double)__bfloat162float(hA[arow * 64 + c]) * __bfloat162float(hB[n * 64 + c]);
        maxerr = fmax(maxerr, fabs(ref - h[m * 64 + n]));
      }
      printf("mode %d (%s) row offset %2d: max err %.3f %s\n", mode, mode == 1 ? "base_offset set" : mode == 2 ? "SBO 1280 patch" : "base_offset 0", off, maxerr, maxerr < 1e-3 ? "OK" : "WRONG");
    }
  return 0;
}
