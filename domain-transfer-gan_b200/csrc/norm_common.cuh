// Helpers shared by the normalisation kernels (norm.cu, norm_fused.cu): 16-byte channel vectors of an NHWC
// plane, plane addressing, activation gradient and the fused gradient load g = (fold(dy) + dy2) * act'(y).
#pragma once
#include "common.cuh"

namespace dtg {

template <typename T>
struct Vec;
template <>
struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static __forceinline__ void load(const void* p, float (&f)[8]) {
    unpack(*reinterpret_cast<const uint4*>(p), f);
  }
  __device__ static __forceinline__ void unpack(const uint4& u, float (&f)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
  __device__ static __forceinline__ void store(void* p, const float (&f)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};
template <>
struct Vec<float> {
  static constexpr int N = 4;
  __device__ static __forceinline__ void load(const void* p, float (&f)[4]) {
    unpack(*reinterpret_cast<const uint4*>(p), f);
  }
  __device__ static __forceinline__ void unpack(const uint4& u, float (&f)[4]) {
    f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y); f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
  }
  __device__ static __forceinline__ void store(void* p, const float (&f)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(round_tf32(f[0]), round_tf32(f[1]), round_tf32(f[2]), round_tf32(f[3]));
  }
};

__device__ __forceinline__ size_t plane_pix(const dtg_plane& p, int n, int y, int x) {
  return (static_cast<size_t>(n) * (p.h + 2 * p.halo) + y + p.halo) * (p.w + 2 * p.halo) + x + p.halo;
}

__device__ __forceinline__ float act_grad(float y, int act) {
  if (act == DTG_ACT_RELU) return y > 0.f ? 1.f : 0.f;
  if (act == DTG_ACT_LRELU) return y > 0.f ? 1.f : 0.2f;
  return 1.f;
}

// g = (fold(dy) + dy2) * act'(y) for one pixel / channel vector
template <typename T>
__device__ __forceinline__ void load_g(const dtg_plane& dy, const dtg_plane& dy2, const dtg_plane& yp, int act, int n,
                                       int y, int x, int c, float (&g)[Vec<T>::N]) {
  constexpr int V = Vec<T>::N;
  const int es = sizeof(T);
  int hts[3], wts[3];
  const int nh = reflect_targets(y, dy.h, dy.halo, hts), nw = reflect_targets(x, dy.w, dy.halo, wts);
#pragma unroll
  for (int i = 0; i < V; ++i) g[i] = 0.f;
  for (int a = 0; a < nh; ++a)
    for (int q = 0; q < nw; ++q) {
      float t[V];
      Vec<T>::load(reinterpret_cast<const uint8_t*>(dy.ptr) + (plane_pix(dy, n, hts[a], wts[q]) * dy.c + c) * es, t);
#pragma unroll
      for (int i = 0; i < V; ++i) g[i] += t[i];
    }
  if (dy2.ptr) {
    float t[V];
    Vec<T>::load(reinterpret_cast<const uint8_t*>(dy2.ptr) + (plane_pix(dy2, n, y, x) * dy2.c + c) * es, t);
#pragma unroll
    for (int i = 0; i < V; ++i) g[i] += t[i];
  }
  if (act != DTG_ACT_NONE) {
    float t[V];
    Vec<T>::load(reinterpret_cast<const uint8_t*>(yp.ptr) + (plane_pix(yp, n, y, x) * yp.c + c) * es, t);
#pragma unroll
    for (int i = 0; i < V; ++i) g[i] *= act_grad(t[i], act);
  }
}


// fused-kernel launchers (norm_fused.cu); return 1 when the geometry is not handled (caller uses the 3-kernel path)
int try_norm_fwd_fused(const dtg_norm_args* a, const dtg_plane* x, const dtg_plane* residual, const float* gamma,
                       const float* beta, float* stats, const dtg_plane* out, cudaStream_t stream);
int try_norm_bwd_fused(const dtg_norm_args* a, const dtg_plane* dy, const dtg_plane* dy2, const dtg_plane* y,
                       const dtg_plane* x, const float* stats, const float* gamma, float* sums, const dtg_plane* dx,
                       const dtg_plane* d_res, cudaStream_t stream);

// TMA-staged persistent-cluster forward / backward (norm_tma.cu); same return convention
int try_norm_fwd_tma(const dtg_norm_args* a, const dtg_plane* x, const dtg_plane* residual, const float* gamma,
                     const float* beta, float* stats, const dtg_plane* out, cudaStream_t stream);
int try_norm_bwd_tma(const dtg_norm_args* a, const dtg_plane* dy, const dtg_plane* dy2, const dtg_plane* y,
                     const dtg_plane* x, const float* stats, const float* gamma, float* sums, const dtg_plane* dx,
                     const dtg_plane* d_res, cudaStream_t stream);

}  // namespace dtg
