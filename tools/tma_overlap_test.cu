// Experiment: is a tiled tensor map with OVERLAPPING rows legal (dim-1 stride 16 B < dim-0 extent 128 B)?  It would give
// a "kw-folded" view of an 8-channel NHWC plane: row(pixel p) = channels of pixels p..p+7, with no materialised im2col.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>
#include "../domain-transfer-gan_b200/csrc/common.cuh"
using namespace dtg;

__global__ void k(const __grid_constant__ CUtensorMap tm, __nv_bfloat16* out, int c1, int c2) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~uintptr_t(1023));
  uint64_t* bar = (uint64_t*)(smem + 16384);
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  __syncthreads();
  if (threadIdx.x == 0) { mbar_expect_tx(bar, 16 * 4 * 128); tma_load_4d(smem, &tm, bar, 0, c1, c2, 0); }
  mbar_wait(bar, 0);
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) out[i] = ((__nv_bfloat16*)smem)[i];
}

int main() {
  const int W = 40, H = 10;
  std::vector<__nv_bfloat16> h(W * H * 8);
  for (int i = 0; i < W * H * 8; ++i) h[i] = __float2bfloat16((float)(i % 251));
  __nv_bfloat16 *d, *o; cudaMalloc(&d, h.size() * 2 + 256); cudaMalloc(&o, 64 * 64 * 2);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tm;
  uint64_t dims[4] = {64, (uint64_t)W, (uint64_t)H, 1};
  uint64_t str[3] = {16, (uint64_t)W * 16, (uint64_t)W * H * 16};
  uint32_t box[4] = {64, 16, 4, 1};
  if (encode_tiled(&tm, DTG_BF16, 4, d, dims, str, box, 0)) { char b[256]; dtg_last_error(b, 256); printf("encode failed: %s\n", b); return 1; }
  k<<<1, 128, 32 * 1024>>>(tm, o, 3, 2);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<__nv_bfloat16> r(64 * 64); cudaMemcpy(r.data(), o, r.size() * 2, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int row = 0; row < 64; ++row) {           // row = (hh*16 + ww): pixel (2+hh, 3+ww); element e -> source pixel px + e/8, ch e%8
    int hh = row / 16, ww = row % 16;
    for (int el = 0; el < 64; ++el) {
      int px = 3 + ww + el / 8, py = 2 + hh;
      float want = 0.f;
      long idx = ((long)py * W + px) * 8 + el % 8;
      if (idx < (long)W * H * 8) want = (float)(idx % 251);   // rows run on into the next image row (flat memory), only the end is OOB
      float got = __bfloat162float(r[row * 64 + el]);
      if (got != want && bad < 5) { printf("row %d el %d got %.0f want %.0f\n", row, el, got, want); }
      bad += got != want;
    }
  }
  printf("overlapping-stride TMA view: %s (%d mismatches)\n", bad ? "MISMATCH" : "OK", bad);
  return 0;
}
