// Single-output-channel head convolutions (stride 1): the last layers of the PatchGAN discriminators
// (networks.py:337 Conv2d(256, 1, 4, padding=1); networks.py:381 Conv2d(128, 1, 4)) and their gradients.
//
// With one output channel the "GEMM" is a matrix-vector product: 2 FLOP per activation byte, i.e. HBM bound by two orders
// of magnitude.  On the tensor-core path it needs N padded 1 -> 16, a 200 KB / 512-column CTA on every SM and ~45 us
// per launch while moving 16 MB; here every activation byte is read once by plain coalesced loads:
//   fwd  : one warp per output pixel, lanes over channels, the filter in shared memory, warp-shuffle reduction
//   dgrad: one warp per input pixel, dx[c] = sum_taps dy[pixel - tap] * w[tap][c]
//   wgrad: one thread per channel, a block per pixel range, per-block partials + fixed-order reduction (deterministic)
// fp32 master weights are used directly (no packed operand).
#include <algorithm>

#include "common.cuh"
#include "norm_common.cuh"

namespace dtg {

constexpr int kH1Threads = 256;
constexpr int kH1MaxTaps = 16;

// filter [cin][taps] (PyTorch [1][cin][kh][kw]) -> shared [taps][cin]
__device__ __forceinline__ void load_filter(const float* __restrict__ w, int cin, int taps, float* wsm) {
  for (int i = threadIdx.x; i < cin * taps; i += blockDim.x) {
    const int c = i / taps, t = i - c * taps;
    wsm[t * cin + c] = w[i];
  }
  __syncthreads();
}

// dot of V activations with V consecutive filter entries in shared memory (16-byte shared loads: a warp reads one
// contiguous KB per tap, no bank conflicts)
template <int V>
__device__ __forceinline__ float dotv(const float (&f)[V], const float* w) {
  float a = 0.f;
#pragma unroll
  for (int i = 0; i < V; i += 4) {
    const float4 q = *reinterpret_cast<const float4*>(w + i);
    a += f[i] * q.x + f[i + 1] * q.y + f[i + 2] * q.z + f[i + 3] * q.w;
  }
  return a;
}

template <typename T>
__global__ void __launch_bounds__(kH1Threads) head1_fwd_kernel(dtg_plane in, const float* __restrict__ w,
                                                               const float* __restrict__ bias, int cin, int kh, int kw,
                                                               int pad, float* __restrict__ out, int oh, int ow) {
  pdl_enter();
  extern __shared__ __align__(16) float wsm[];
  constexpr int V = Vec<T>::N;
  const int taps = kh * kw;
  load_filter(w, cin, taps, wsm);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const unsigned total = static_cast<unsigned>(in.n) * oh * ow;
  const float b0 = bias ? bias[0] : 0.f;
  for (unsigned o = blockIdx.x * wpb + warp; o < total; o += gridDim.x * wpb) {
    const int ox = o % ow;
    const int oy = (o / ow) % oh;
    const int n = o / (static_cast<unsigned>(ow) * oh);
    float acc = 0.f;
    int t = 0;
    for (int a = 0; a < kh; ++a) {
      const int iy = oy + a - pad;
      for (int q = 0; q < kw; ++q, ++t) {
        const int ix = ox + q - pad;
        if (iy < 0 || iy >= in.h || ix < 0 || ix >= in.w) continue;       // zero padding
        const uint8_t* px = reinterpret_cast<const uint8_t*>(in.ptr) + plane_pix(in, n, iy, ix) * in.c * sizeof(T);
        for (int c = lane * V; c < cin; c += 32 * V) {
          float f[V];
          Vec<T>::load(px + c * sizeof(T), f);
          acc += dotv<V>(f, wsm + t * cin + c);
        }
      }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) out[o] = acc + b0;
  }
}

template <typename T>
__device__ __forceinline__ float ld_ch0(const dtg_plane& p, size_t pix) {
  if constexpr (sizeof(T) == 2)
    return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.ptr)[pix * p.c]);
  else
    return reinterpret_cast<const float*>(p.ptr)[pix * p.c];
}

// lane t (< taps) fetches the seed gradient that filter tap t pairs with input pixel (n, y, x); 0 outside
template <typename T>
__device__ __forceinline__ float tap_grad(const dtg_plane& dy, int n, int y, int x, int kw, int pad, int taps, int lane) {
  if (lane >= taps) return 0.f;
  const int a = lane / kw, q = lane - a * kw;
  const int oy = y + pad - a, ox = x + pad - q;
  if (oy < 0 || oy >= dy.h || ox < 0 || ox >= dy.w) return 0.f;
  return ld_ch0<T>(dy, plane_pix(dy, n, oy, ox));
}

template <typename T>
__global__ void __launch_bounds__(kH1Threads) head1_dgrad_kernel(dtg_plane dy, const float* __restrict__ w, int cin, int kh,
                                                                 int kw, int pad, dtg_plane dx) {
  pdl_enter();
  extern __shared__ __align__(16) float wsm[];
  constexpr int V = Vec<T>::N;
  const int taps = kh * kw;
  load_filter(w, cin, taps, wsm);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const unsigned total = static_cast<unsigned>(dx.n) * dx.h * dx.w;
  for (unsigned o = blockIdx.x * wpb + warp; o < total; o += gridDim.x * wpb) {
    const int x = o % dx.w;
    const int y = (o / dx.w) % dx.h;
    const int n = o / (static_cast<unsigned>(dx.w) * dx.h);
    const float gl = tap_grad<T>(dy, n, y, x, kw, pad, taps, lane);
    uint8_t* px = reinterpret_cast<uint8_t*>(dx.ptr) + plane_pix(dx, n, y, x) * dx.c * sizeof(T);
    for (int c0 = 0; c0 < dx.c; c0 += 32 * V) {           // warp-uniform trip count: the shuffles need every lane
      const int c = c0 + lane * V;
      float f[V];
#pragma unroll
      for (int i = 0; i < V; ++i) f[i] = 0.f;
      for (int t = 0; t < taps; ++t) {
        const float g = __shfl_sync(0xffffffffu, gl, t);
        if (c + V <= cin) {
          const float* wv = wsm + t * cin + c;
#pragma unroll
          for (int i = 0; i < V; i += 4) {
            const float4 q = *reinterpret_cast<const float4*>(wv + i);
            f[i] += g * q.x;
            f[i + 1] += g * q.y;
            f[i + 2] += g * q.z;
            f[i + 3] += g * q.w;
          }
        }
      }
      if (c < dx.c) Vec<T>::store(px + c * sizeof(T), f);
    }
  }
}

// Weight gradient.  A block = 8 warps = (pixel group) x (64-channel quarter): lane owns 2 channels, loops over its
// group's input pixels (each activation is read once), lanes 0..taps-1 fetch the seed gradients that pair with the pixel
// and broadcast them by shuffle.  Groups are summed through shared memory in index order, blocks by the reduce kernel.
template <typename T>
__global__ void __launch_bounds__(kH1Threads) head1_wgrad_kernel(dtg_plane dy, dtg_plane in, int cin, int kh, int kw, int pad,
                                                                 float* __restrict__ part, int pix_per_block) {
  pdl_enter();
  __shared__ float red[8 * kH1MaxTaps * 64];          // [warp][tap][64 channels]
  const int taps = kh * kw;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nq = (cin + 63) / 64, ng = 8 / nq;        // channel quarters, pixel groups (cin <= 512)
  const int q = warp % nq, pg = warp / nq;
  const unsigned total = static_cast<unsigned>(in.n) * in.h * in.w;
  const unsigned p0 = blockIdx.x * static_cast<unsigned>(pix_per_block), p1 = min(total, p0 + pix_per_block);
  const int c = q * 64 + lane * 2;
  float acc0[kH1MaxTaps], acc1[kH1MaxTaps];
#pragma unroll
  for (int t = 0; t < kH1MaxTaps; ++t) acc0[t] = acc1[t] = 0.f;
  if (pg < ng) {
    for (unsigned p = p0 + pg; p < p1; p += ng) {
      const int x = p % in.w;
      const int y = (p / in.w) % in.h;
      const int n = p / (static_cast<unsigned>(in.w) * in.h);
      float v0 = 0.f, v1 = 0.f;
      if (c + 1 < in.c) {
        const size_t e = plane_pix(in, n, y, x) * in.c + c;
        if constexpr (sizeof(T) == 2) {
          const uint32_t u = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const __nv_bfloat16*>(in.ptr) + e);
          v0 = __uint_as_float(u << 16);
          v1 = __uint_as_float(u & 0xFFFF0000u);
        } else {
          const float2 u = *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(in.ptr) + e);
          v0 = u.x;
          v1 = u.y;
        }
      }
      const float gl = tap_grad<T>(dy, n, y, x, kw, pad, taps, lane);
#pragma unroll
      for (int t = 0; t < kH1MaxTaps; ++t) {
        const float g = __shfl_sync(0xffffffffu, gl, t);     // 0 for t >= taps
        acc0[t] += g * v0;
        acc1[t] += g * v1;
      }
    }
  }
#pragma unroll
  for (int t = 0; t < kH1MaxTaps; ++t) {
    red[(warp * kH1MaxTaps + t) * 64 + lane * 2] = acc0[t];
    red[(warp * kH1MaxTaps + t) * 64 + lane * 2 + 1] = acc1[t];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < taps * cin; i += blockDim.x) {
    const int t = i / cin, ch = i - t * cin;
    const int qq = ch / 64, cl = ch - qq * 64;
    float s = 0.f;
    for (int g = 0; g < ng; ++g) s += red[((g * nq + qq) * kH1MaxTaps + t) * 64 + cl];
    part[static_cast<size_t>(blockIdx.x) * taps * cin + i] = s;
  }
}

// dw[c][tap] += sum_b part[b][tap][c], blocks summed in index order (deterministic)
__global__ void __launch_bounds__(256) head1_wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ dw, int nblocks,
                                                                 int taps, int cin) {
  pdl_enter();
  const int total = taps * cin;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < nblocks; ++b) s += part[static_cast<size_t>(b) * total + i];
    const int t = i / cin, c = i - t * cin;
    dw[c * taps + t] += s;
  }
}

static int h1_grid(size_t warps_needed) {
  return static_cast<int>(std::max<size_t>(1, std::min<size_t>((warps_needed + 7) / 8, 148 * 8)));
}

static int h1_check(const dtg_plane* a, int cin, int kh, int kw, const char* what) {
  DTG_REQUIRE(a && a->ptr, "%s: null plane", what);
  DTG_REQUIRE(kh >= 1 && kw >= 1 && kh * kw <= kH1MaxTaps, "%s: at most %d taps", what, kH1MaxTaps);
  DTG_REQUIRE(cin >= 8 && cin % 8 == 0 && cin <= 512 && cin <= a->c && cin * kh * kw * sizeof(float) <= 48 * 1024,
              "%s: cin %d unsupported (multiple of 8, <= 512)", what, cin);
  DTG_REQUIRE((a->c * elem_size(a->dtype)) % 16 == 0, "%s: channel pitch must be a multiple of 16 bytes", what);
  return DTG_OK;
}

constexpr int kH1WgradBlocks = 296;

}  // namespace dtg

using namespace dtg;

extern "C" int dtg_head1_fwd(const dtg_plane* in, const float* w, const float* bias, int cin, int kh, int kw, int pad,
                             float* out_nchw, int oh, int ow, void* stream) {
  int rc = h1_check(in, cin, kh, kw, "dtg_head1_fwd");
  if (rc != DTG_OK) return rc;
  DTG_REQUIRE(w && out_nchw && in->halo == 0, "dtg_head1_fwd: null argument or haloed input");
  DTG_REQUIRE(in->h + 2 * pad - kh + 1 == oh && in->w + 2 * pad - kw + 1 == ow, "dtg_head1_fwd: output extent mismatch");
  const size_t smem = static_cast<size_t>(cin) * kh * kw * sizeof(float);
  const int grid = h1_grid(static_cast<size_t>(in->n) * oh * ow);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (in->dtype == DTG_F32)
    DTG_CHECK_CUDA(launch_k(head1_fwd_kernel<float>, grid, kH1Threads, smem, s, *in, w, bias, cin, kh, kw, pad, out_nchw, oh, ow));
  else
    DTG_CHECK_CUDA(launch_k(head1_fwd_kernel<__nv_bfloat16>, grid, kH1Threads, smem, s, *in, w, bias, cin, kh, kw, pad, out_nchw, oh, ow));
  return DTG_OK;
}

extern "C" int dtg_head1_dgrad(const dtg_plane* dy, const float* w, int cin, int kh, int kw, int pad, const dtg_plane* dx,
                               void* stream) {
  int rc = h1_check(dx, cin, kh, kw, "dtg_head1_dgrad");
  if (rc != DTG_OK) return rc;
  DTG_REQUIRE(dy && dy->ptr && w && dy->dtype == dx->dtype && dy->n == dx->n && dx->halo == 0 && dy->halo == 0,
              "dtg_head1_dgrad: bad planes");
  DTG_REQUIRE(dx->h + 2 * pad - kh + 1 == dy->h && dx->w + 2 * pad - kw + 1 == dy->w, "dtg_head1_dgrad: extent mismatch");
  const size_t smem = static_cast<size_t>(cin) * kh * kw * sizeof(float);
  const int grid = h1_grid(static_cast<size_t>(dx->n) * dx->h * dx->w);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dx->dtype == DTG_F32)
    DTG_CHECK_CUDA(launch_k(head1_dgrad_kernel<float>, grid, kH1Threads, smem, s, *dy, w, cin, kh, kw, pad, *dx));
  else
    DTG_CHECK_CUDA(launch_k(head1_dgrad_kernel<__nv_bfloat16>, grid, kH1Threads, smem, s, *dy, w, cin, kh, kw, pad, *dx));
  return DTG_OK;
}

extern "C" size_t dtg_head1_wgrad_workspace_bytes(int cin, int kh, int kw) {
  return static_cast<size_t>(kH1WgradBlocks) * cin * kh * kw * sizeof(float);
}

extern "C" int dtg_head1_wgrad(const dtg_plane* dy, const dtg_plane* in, float* dw, int cin, int kh, int kw, int pad,
                               void* workspace, size_t workspace_bytes, void* stream) {
  int rc = h1_check(in, cin, kh, kw, "dtg_head1_wgrad");
  if (rc != DTG_OK) return rc;
  DTG_REQUIRE(dy && dy->ptr && dw && workspace && dy->dtype == in->dtype && dy->n == in->n && in->halo == 0 && dy->halo == 0,
              "dtg_head1_wgrad: bad arguments");
  DTG_REQUIRE(in->h + 2 * pad - kh + 1 == dy->h && in->w + 2 * pad - kw + 1 == dy->w, "dtg_head1_wgrad: extent mismatch");
  DTG_REQUIRE(workspace_bytes >= dtg_head1_wgrad_workspace_bytes(cin, kh, kw), "dtg_head1_wgrad: workspace too small");
  const size_t total = static_cast<size_t>(in->n) * in->h * in->w;
  const int nblocks = static_cast<int>(std::min<size_t>(kH1WgradBlocks, total));
  const int ppb = static_cast<int>((total + nblocks - 1) / nblocks);
  const int used = static_cast<int>((total + ppb - 1) / ppb);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* part = reinterpret_cast<float*>(workspace);
  if (in->dtype == DTG_F32)
    DTG_CHECK_CUDA(launch_k(head1_wgrad_kernel<float>, used, kH1Threads, 0, s, *dy, *in, cin, kh, kw, pad, part, ppb));
  else
    DTG_CHECK_CUDA(launch_k(head1_wgrad_kernel<__nv_bfloat16>, used, kH1Threads, 0, s, *dy, *in, cin, kh, kw, pad, part, ppb));
  const int tot = cin * kh * kw;
  DTG_CHECK_CUDA(launch_k(head1_wgrad_reduce_kernel, std::max(1, std::min((tot + 255) / 256, 64)), 256, 0, s, part, dw, used, kh * kw, cin));
  return DTG_OK;
}
