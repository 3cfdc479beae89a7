// Experiment: TMA tiled-load cost as a function of the box shape, row width, swizzle and out-of-bounds rows.
// One CTA per SM; one thread keeps 4 boxes in flight over a ring of smem stages; reports clk per box and per row.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../domain-transfer-gan_b200/csrc/common.cuh"
using namespace dtg;

struct Cfg { int c0, c1, c2, step1, step2, n1, n2, bytes, iters, store; };

__global__ void __launch_bounds__(64, 1) k(const __grid_constant__ CUtensorMap tm, Cfg c, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~uintptr_t(1023));
  constexpr int S = 4, STAGE = 40 * 1024;
  uint64_t* bar = (uint64_t*)(smem + S * STAGE);
  if (threadIdx.x == 0) { for (int i = 0; i < S; ++i) mbar_init(&bar[i], 1); mbar_fence_init(); }
  __syncthreads();
  long long t0 = 0, t1 = 0;
  if (threadIdx.x == 0) {
    const int img = blockIdx.x % 80;
    t0 = clock64();
    if (!c.store) {
      for (int i = 0; i < c.iters + S; ++i) {
        const int s = i % S;
        if (i >= S) mbar_wait(&bar[s], ((i / S) - 1) & 1);
        if (i < c.iters) {
          const int a = (i * 7 + blockIdx.x) % c.n1, b = (i * 3 + blockIdx.x) % c.n2;
          mbar_expect_tx(&bar[s], c.bytes);
          tma_load_4d(smem + s * STAGE, &tm, &bar[s], c.c0, c.c1 + a * c.step1, c.c2 + b * c.step2, img);
        }
      }
    } else {
      for (int i = 0; i < c.iters; ++i) {
        const int a = (i * 7 + blockIdx.x) % c.n1, b = (i * 3 + blockIdx.x) % c.n2;
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                     ::"l"((uint64_t)&tm), "r"(smem_u32(smem + (i % S) * STAGE)), "r"(c.c0), "r"(c.c1 + a * c.step1),
                     "r"(c.c2 + b * c.step2), "r"(img) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    t1 = clock64();
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

int main() {
  const int N = 80, H = 66, W = 66;
  void* buf; cudaMalloc(&buf, (size_t)N * H * W * 128); cudaMemset(buf, 0, (size_t)N * H * W * 128);
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct T { const char* name; int C; int bc, bw, bh; int sw; int c1, c2, step1, step2, n1, n2; int store; } tests[] = {
      {"c64 box{64,32,4}  aligned        ", 64, 64, 32, 4, 1, 0, 0, 32, 4, 2, 15, 0},
      {"c64 box{64,8,16}  aligned        ", 64, 64, 8, 16, 1, 0, 0, 8, 16, 8, 4, 0},
      {"c64 box{64,10,18} interior (+7)  ", 64, 64, 10, 18, 1, 7, 15, 8, 16, 6, 2, 0},
      {"c64 box{64,10,18} w/ OOB (-1)    ", 64, 64, 10, 18, 1, -1, -1, 8, 16, 8, 4, 0},
      {"c64 box{64,14,22} interior       ", 64, 64, 14, 22, 1, 5, 13, 8, 16, 6, 2, 0},
      {"c64 box{64,180,1} one long row   ", 64, 64, 180, 1, 1, 0, 0, 0, 1, 1, 60, 0},
      {"c64 box{64,60,3}                 ", 64, 64, 60, 3, 1, 0, 0, 0, 3, 1, 20, 0},
      {"c32 box{32,10,18} sw64 interior  ", 32, 32, 10, 18, 3, 7, 15, 8, 16, 6, 2, 0},
      {"c16 box{16,14,22} sw32 interior  ", 16, 16, 14, 22, 4, 5, 13, 8, 16, 6, 2, 0},
      {"c16 box{16,14,22} sw32 OOB       ", 16, 16, 14, 22, 4, -3, -3, 8, 16, 8, 4, 0},
      {"c64 STORE box{64,8,4}            ", 64, 64, 8, 4, 1, 0, 0, 8, 4, 8, 16, 1},
      {"c64 STORE box{64,8,16}           ", 64, 64, 8, 16, 1, 0, 0, 8, 16, 8, 4, 1},
      {"c32 STORE box{32,8,4} sw64       ", 32, 32, 8, 4, 3, 0, 0, 8, 4, 8, 16, 1},
  };
  for (auto& t : tests) {
    CUtensorMap tm;
    uint64_t dims[4] = {(uint64_t)t.C, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t str[3] = {(uint64_t)t.C * 2, (uint64_t)W * t.C * 2, (uint64_t)H * W * t.C * 2};
    uint32_t box[4] = {(uint32_t)t.bc, (uint32_t)t.bw, (uint32_t)t.bh, 1};
    if (encode_tiled(&tm, DTG_BF16, 4, buf, dims, str, box, t.sw)) { printf("%s encode failed\n", t.name); continue; }
    Cfg c{0, t.c1, t.c2, t.step1, t.step2, t.n1, t.n2, t.bc * 2 * t.bw * t.bh, 400, t.store};
    k<<<148, 64, 170 * 1024>>>(tm, c, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", t.name, cudaGetErrorString(e)); return 1; }
    long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    const double per = (double)mx / c.iters;
    printf("%s: %7.0f clk/box  %5.1f clk/row  %5.1f B/clk/SM\n", t.name, per, per / (t.bw * t.bh), c.bytes / per);
  }
  return 0;
}
