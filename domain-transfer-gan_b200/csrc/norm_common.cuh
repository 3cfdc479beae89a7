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


// pixel cursor over a CTA's pixel range that tracks the pixel offsets inside up to three (possibly haloed) planes
// incrementally (no per-pixel divisions)
struct PixCur {
  int p, py, px;
  int o0, o1;    // pixel offsets inside planes with halo h0/h1: (py + h) * (w + 2h) + px + h
  __device__ __forceinline__ void init(int p0, int w, int h0, int h1) {
    p = p0;
    py = p0 / w;
    px = p0 - py * w;
    o0 = (py + h0) * (w + 2 * h0) + px + h0;
    o1 = (py + h1) * (w + 2 * h1) + px + h1;
  }
  // step = q * w + r (q, r precomputed): at most one wrap test per step
  template <bool HAL>
  __device__ __forceinline__ void advance(int step, int q, int r, int w, int h0, int h1) {
    p += step;
    if (HAL) {
      px += r;
      int rows = q;
      if (px >= w) {
        px -= w;
        ++rows;
      }
      py += rows;
      o0 += step + 2 * h0 * rows;
      o1 += step + 2 * h1 * rows;
    } else {
      o0 += step;
      o1 += step;
    }
  }
};

// g = (fold(dy) + dy2) * act'(y) with precomputed plane offsets; the reflect fold only touches border pixels
template <typename T, int ACT, bool HAL>
__device__ __forceinline__ void load_g_fast(const dtg_plane& dy, const uint8_t* dyb, const uint8_t* dy2b, const uint8_t* yb,
                                            const PixCur& it, int n, int c, float (&g)[Vec<T>::N]) {
  constexpr int V = Vec<T>::N;
  const size_t es = sizeof(T);
  Vec<T>::load(dyb + static_cast<size_t>(it.o0) * dy.c * es, g);
  if (HAL && dy.halo > 0) {
    const int hl = dy.halo;
    if (hl == 1 && dy.h >= 4 && dy.w >= 4) {
      // reflection-pad(1) backward: pixel 1 also receives halo -1 (offset -2), pixel L-2 receives halo L (offset +2)
      const int wb = dy.w + 2;
      const int dr = it.py == 1 ? -2 * wb : (it.py == dy.h - 2 ? 2 * wb : 0);
      const int dc = it.px == 1 ? -2 : (it.px == dy.w - 2 ? 2 : 0);
      if (dr != 0) {
        float t[V];
        Vec<T>::load(dyb + static_cast<size_t>(it.o0 + dr) * dy.c * es, t);
#pragma unroll
        for (int i = 0; i < V; ++i) g[i] += t[i];
      }
      if (dc != 0) {
        float t[V];
        Vec<T>::load(dyb + static_cast<size_t>(it.o0 + dc) * dy.c * es, t);
#pragma unroll
        for (int i = 0; i < V; ++i) g[i] += t[i];
        if (dr != 0) {
          Vec<T>::load(dyb + static_cast<size_t>(it.o0 + dr + dc) * dy.c * es, t);
#pragma unroll
          for (int i = 0; i < V; ++i) g[i] += t[i];
        }
      }
    } else {
      const bool border = (it.py >= 1 && it.py <= hl) || (it.py <= dy.h - 2 && it.py >= dy.h - 1 - hl) ||
                          (it.px >= 1 && it.px <= hl) || (it.px <= dy.w - 2 && it.px >= dy.w - 1 - hl);
      if (border) {
        int hts[3], wts[3];
        const int nh = reflect_targets(it.py, dy.h, hl, hts), nw = reflect_targets(it.px, dy.w, hl, wts);
        for (int a = 0; a < nh; ++a)
          for (int q = 0; q < nw; ++q) {
            if (a + q == 0) continue;
            float t[V];
            Vec<T>::load(reinterpret_cast<const uint8_t*>(dy.ptr) + (plane_pix(dy, n, hts[a], wts[q]) * dy.c + c) * es, t);
#pragma unroll
            for (int i = 0; i < V; ++i) g[i] += t[i];
          }
      }
    }
  }
  if (dy2b) {
    float t[V];
    Vec<T>::load(dy2b + static_cast<size_t>(it.p) * dy.c * es, t);
#pragma unroll
    for (int i = 0; i < V; ++i) g[i] += t[i];
  }
  if (ACT != DTG_ACT_NONE) {
    float t[V];
    Vec<T>::load(yb + static_cast<size_t>(it.o1) * dy.c * es, t);
#pragma unroll
    for (int i = 0; i < V; ++i) g[i] = t[i] > 0.f ? g[i] : (ACT == DTG_ACT_LRELU ? 0.2f * g[i] : 0.f);
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

// two-phase streaming kernels (norm_lean.cu); same return convention; `partial`: n * 32 * c float2 of scratch
int try_norm_fwd_lean(const dtg_norm_args* a, const dtg_plane* x, const dtg_plane* residual, const float* gamma,
                      const float* beta, float* stats, float* partial, const dtg_plane* out, cudaStream_t stream);
int try_norm_bwd_lean(const dtg_norm_args* a, const dtg_plane* dy, const dtg_plane* dy2, const dtg_plane* y,
                      const dtg_plane* x, const float* stats, const float* gamma, float* sums, float* partial,
                      const dtg_plane* dx, const dtg_plane* d_res, cudaStream_t stream);

}  // namespace dtg
