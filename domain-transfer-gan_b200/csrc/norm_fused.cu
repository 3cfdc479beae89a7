// Single-launch (conditional) instance normalisation, forward and backward, on thread-block clusters.
//
// A "slab" is one sample x one 128-byte channel block x all pixels.  A cluster of CS CTAs owns a slab and
// splits its pixels: pass 1 streams the slab once from HBM and reduces the per-channel statistics
// (block reduction in fixed order -> per-CTA partials in shared memory -> every CTA sums the CS partials of its
// peers through distributed shared memory, in rank order: deterministic, no atomics, no global round trip);
// pass 2 re-reads the CTA's own pixels -- a few hundred KB that are still L2-resident -- and writes the result.
// HBM traffic is therefore one read of the inputs and one write of the outputs, instead of the two full reads
// plus three finalisation launches of the split path in norm.cu (which remains for batch norm and odd shapes).
// Formulas: SURVEY.md 9.1 (modules.py:83-97,120-132 and their autograd).
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "norm_common.cuh"

namespace cg = cooperative_groups;

namespace dtg {

constexpr int kFT = 256;          // threads per CTA
constexpr int kSlabCh = 64;       // max channels of a slab (128 B of bf16)

constexpr int kMaxCluster = 8;

// Fixed-order block reduction of per-thread (s1[V], s2[V]) over the pixel lanes, then PUSH: the per-channel pair
// is stored into slot `rank` of every cluster peer's allpart[] through distributed shared memory.  After ONE
// cluster barrier every CTA owns all partials locally and sums them in rank order (deterministic); nobody reads
// remote shared memory afterwards, so CTAs may exit independently.
template <int V, int NT = kFT>
__device__ __forceinline__ void block_reduce_push(const float (&s1)[V], const float (&s2)[V], int nv, float* red,
                                                  float (*allpart)[kSlabCh * 2], cg::cluster_group& cl, int cs, int rank) {
  const int t = threadIdx.x;
  const int lanes = NT / nv;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    red[(t * V + i) * 2] = s1[i];
    red[(t * V + i) * 2 + 1] = s2[i];
  }
  __syncthreads();
  if (t < nv * V) {
    const int v = t / V, i = t % V;
    float a = 0.f, b = 0.f;
#pragma unroll 8
    for (int l = 0; l < lanes; ++l) {
      const int src = l * nv + v;
      const float2 x = *reinterpret_cast<const float2*>(red + (src * V + i) * 2);
      a += x.x;
      b += x.y;
    }
    float2* mine = reinterpret_cast<float2*>(&allpart[rank][t * 2]);
    for (int r = 0; r < cs; ++r) *cl.map_shared_rank(mine, r) = make_float2(a, b);
  }
}

// sum of the cluster's partials for channel t (call after the cluster barrier)
__device__ __forceinline__ float2 cluster_total(float (*allpart)[kSlabCh * 2], int cs, int t) {
  float a = 0.f, b = 0.f;
  for (int r = 0; r < cs; ++r) {
    a += allpart[r][t * 2];
    b += allpart[r][t * 2 + 1];
  }
  return make_float2(a, b);
}

struct PixIter {
  int p, py, px;
  __device__ __forceinline__ void init(int p0, int w) {
    p = p0;
    py = p0 / w;
    px = p0 - py * w;
  }
  __device__ __forceinline__ void advance(int step, int w) {
    p += step;
    px += step;
    while (px >= w) {
      px -= w;
      ++py;
    }
  }
};

template <typename T>
__global__ void __launch_bounds__(kFT) norm_fwd_fused_kernel(dtg_plane x, dtg_plane res, dtg_plane out,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float* __restrict__ stats,
                                                             int mode, int act, float eps, int nv, int chunk) {
  pdl_enter();
  constexpr int V = Vec<T>::N;
  __shared__ float red[kFT * V * 2];
  __shared__ float allpart[kMaxCluster][kSlabCh * 2];
  __shared__ float2 coef[kSlabCh];
  cg::cluster_group cl = cg::this_cluster();
  const int cs = cl.num_blocks(), rank = cl.block_rank();
  const int slab_ch = nv * V;
  const int v = threadIdx.x % nv, lane = threadIdx.x / nv, lanes = kFT / nv;
  const int n = blockIdx.y;
  const int c = blockIdx.x * slab_ch + v * V;
  const int hw = x.h * x.w;
  const int p0 = rank * chunk, p1 = min(hw, p0 + chunk);
  const uint8_t* xb = reinterpret_cast<const uint8_t*>(x.ptr) + (static_cast<size_t>(n) * hw * x.c + c) * sizeof(T);
  const size_t pitch = static_cast<size_t>(x.c) * sizeof(T);

  // ---- pass 1: shifted sums (K = first pixel of the sample, identical in every CTA of the cluster)
  float K[V], s1[V], s2[V];
  Vec<T>::load(xb, K);
#pragma unroll
  for (int i = 0; i < V; ++i) s1[i] = s2[i] = 0.f;
#pragma unroll 4
  for (int p = p0 + lane; p < p1; p += lanes) {
    float f[V];
    Vec<T>::load(xb + p * pitch, f);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float d = f[i] - K[i];
      s1[i] += d;
      s2[i] += d * d;
    }
  }
  block_reduce_push<V>(s1, s2, nv, red, allpart, cl, cs, rank);
  cl.sync();
  if (threadIdx.x < slab_ch) {
    const int t = threadIdx.x;
    const float2 tot = cluster_total(allpart, cs, t);
    const float a1 = tot.x, a2 = tot.y;
    const int ch = blockIdx.x * slab_ch + t;
    const size_t first = static_cast<size_t>(n) * hw * x.c + ch;
    float Kc;
    if constexpr (sizeof(T) == 2)
      Kc = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x.ptr)[first]);
    else
      Kc = reinterpret_cast<const float*>(x.ptr)[first];
    const float m = static_cast<float>(hw);
    const float mean = Kc + a1 / m;
    const float d = mode == DTG_NORM_COND_INSTANCE ? m - 1.f : m;
    float var = (a2 - a1 * a1 / m) / d;
    var = var < 0.f ? 0.f : var;
    const float rstd = rsqrtf(var + eps);
    const size_t nc = static_cast<size_t>(n) * x.c + ch;
    const float ga = mode == DTG_NORM_COND_INSTANCE ? gamma[nc] : gamma[ch];
    const float be = mode == DTG_NORM_COND_INSTANCE ? beta[nc] : beta[ch];
    const float a = rstd * ga;
    coef[t] = make_float2(a, be - mean * a);
    if (rank == 0) {
      stats[nc * 2] = mean;
      stats[nc * 2 + 1] = rstd;
    }
  }
  __syncthreads();    // coef[] is visible block-wide

  // ---- pass 2: y = act(x*a + b (+ residual)), mirrored into the output halo
  float ca[V], cb[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float2 ab = coef[v * V + i];
    ca[i] = ab.x;
    cb[i] = ab.y;
  }
  PixIter it;
  it.init(p0 + lane, x.w);
#pragma unroll 2
  for (; it.p < p1; it.advance(lanes, x.w)) {
    float f[V];
    Vec<T>::load(xb + it.p * pitch, f);
#pragma unroll
    for (int i = 0; i < V; ++i) f[i] = f[i] * ca[i] + cb[i];
    if (res.ptr) {
      float t[V];
      Vec<T>::load(reinterpret_cast<const uint8_t*>(res.ptr) + (plane_pix(res, n, it.py, it.px) * res.c + c) * sizeof(T), t);
#pragma unroll
      for (int i = 0; i < V; ++i) f[i] += t[i];
    }
#pragma unroll
    for (int i = 0; i < V; ++i) f[i] = apply_act(f[i], act);
    uint8_t* ob = reinterpret_cast<uint8_t*>(out.ptr);
    Vec<T>::store(ob + (plane_pix(out, n, it.py, it.px) * out.c + c) * sizeof(T), f);
    if (out.halo > 0) {
      int hts[3], wts[3];
      const int nh = reflect_targets(it.py, out.h, out.halo, hts), nw = reflect_targets(it.px, out.w, out.halo, wts);
      if (nh * nw > 1)
        for (int a = 0; a < nh; ++a)
          for (int q = 0; q < nw; ++q)
            if (a + q > 0) Vec<T>::store(ob + (plane_pix(out, n, hts[a], wts[q]) * out.c + c) * sizeof(T), f);
    }
  }
}

template <typename T, int ACT, bool HAL>
__global__ void __launch_bounds__(kFT, 3) norm_bwd_fused_kernel(dtg_plane dy, dtg_plane dy2, dtg_plane yp, dtg_plane x,
                                                                const float* __restrict__ stats,
                                                                const float* __restrict__ gamma, float* __restrict__ sums,
                                                                dtg_plane dx, dtg_plane dres, int mode, int nv,
                                                                int chunk) {
  pdl_enter();
  constexpr int V = Vec<T>::N;
  __shared__ float red[kFT * V * 2];
  __shared__ float allpart[kMaxCluster][kSlabCh * 2];
  __shared__ float4 kco[kSlabCh];
  cg::cluster_group cl = cg::this_cluster();
  const int cs = cl.num_blocks(), rank = cl.block_rank();
  const int slab_ch = nv * V;
  const int v = threadIdx.x % nv, lane = threadIdx.x / nv, lanes = kFT / nv;
  const int n = blockIdx.y;
  const int c = blockIdx.x * slab_ch + v * V;
  const int hw = x.h * x.w;
  const int p0 = rank * chunk, p1 = min(hw, p0 + chunk);
  const size_t es = sizeof(T);
  const size_t pitch = static_cast<size_t>(x.c) * es;      // every plane here has x.c channels
  const uint8_t* xb = reinterpret_cast<const uint8_t*>(x.ptr) + (static_cast<size_t>(n) * hw * x.c + c) * es;
  const int hdy = dy.halo, hy = yp.ptr ? yp.halo : 0;
  const uint8_t* dyb = reinterpret_cast<const uint8_t*>(dy.ptr) +
                       (static_cast<size_t>(n) * (dy.h + 2 * hdy) * (dy.w + 2 * hdy) * dy.c + c) * es;
  const uint8_t* yb = yp.ptr ? reinterpret_cast<const uint8_t*>(yp.ptr) +
                                   (static_cast<size_t>(n) * (yp.h + 2 * hy) * (yp.w + 2 * hy) * yp.c + c) * es
                             : nullptr;
  const uint8_t* dy2b = dy2.ptr ? reinterpret_cast<const uint8_t*>(dy2.ptr) + (static_cast<size_t>(n) * hw * dy2.c + c) * es
                                : nullptr;
  uint8_t* dxb = reinterpret_cast<uint8_t*>(dx.ptr) + (static_cast<size_t>(n) * hw * dx.c + c) * es;
  uint8_t* drb = dres.ptr ? reinterpret_cast<uint8_t*>(dres.ptr) + (static_cast<size_t>(n) * hw * dres.c + c) * es : nullptr;

  float mean[V], rstd[V], s1[V], s2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    s1[i] = s2[i] = 0.f;
    const float2 mr = *reinterpret_cast<const float2*>(stats + (static_cast<size_t>(n) * x.c + c + i) * 2);
    mean[i] = mr.x;
    rstd[i] = mr.y;
  }
  // ---- pass 1: A = sum g, B = sum g * xhat
  PixCur it;
  const int adv_q = lanes / x.w, adv_r = lanes - adv_q * x.w;
  it.init(p0 + lane, x.w, hdy, hy);
#pragma unroll 2
  for (; it.p < p1; it.template advance<HAL>(lanes, adv_q, adv_r, x.w, hdy, hy)) {
    float g[V], f[V];
    load_g_fast<T, ACT, HAL>(dy, dyb, dy2b, yb, it, n, c, g);
    Vec<T>::load(xb + it.p * pitch, f);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      s1[i] += g[i];
      s2[i] += g[i] * (f[i] - mean[i]) * rstd[i];
    }
  }
  block_reduce_push<V>(s1, s2, nv, red, allpart, cl, cs, rank);
  cl.sync();
  if (threadIdx.x < slab_ch) {
    const int t = threadIdx.x;
    const float2 tot = cluster_total(allpart, cs, t);
    const float A = tot.x, B = tot.y;
    const int ch = blockIdx.x * slab_ch + t;
    const size_t nc = static_cast<size_t>(n) * x.c + ch;
    const float m = static_cast<float>(hw);
    const float d = mode == DTG_NORM_COND_INSTANCE ? m - 1.f : m;
    const float ga = mode == DTG_NORM_COND_INSTANCE ? gamma[nc] : gamma[ch];
    kco[t] = make_float4(stats[nc * 2 + 1] * ga, A / m, B / d, 0.f);
    if (rank == 0) {
      sums[nc * 2] = A;
      sums[nc * 2 + 1] = B;
    }
  }
  __syncthreads();

  // ---- pass 2 (L2-resident re-read): dx = k0 * (g - kA - xhat * kB); d_res = g
  float k0[V], kA[V], kB[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float4 k = kco[v * V + i];
    k0[i] = k.x;                // dx = k0 * (g - kA) - (x - mean) * (k0 * rstd * B/d)
    kA[i] = k.y;
    kB[i] = k.x * rstd[i] * k.z;
  }
  it.init(p0 + lane, x.w, hdy, hy);
#pragma unroll 2
  for (; it.p < p1; it.template advance<HAL>(lanes, adv_q, adv_r, x.w, hdy, hy)) {
    float g[V], f[V];
    load_g_fast<T, ACT, HAL>(dy, dyb, dy2b, yb, it, n, c, g);
    Vec<T>::load(xb + it.p * pitch, f);
    if (drb) Vec<T>::store(drb + it.p * pitch, g);
#pragma unroll
    for (int i = 0; i < V; ++i) g[i] = k0[i] * (g[i] - kA[i]) - (f[i] - mean[i]) * kB[i];
    Vec<T>::store(dxb + it.p * pitch, g);
  }
}

// Register-resident variant for slabs of <= 8 * lanes * PPT pixels (the 32x32 residual-stack planes): every thread
// issues ALL its loads up front (PPT pixels x {dy, dy2, y, x}), keeps g and xhat in registers across the cluster
// reduction, and pass 2 is pure arithmetic + stores: one HBM read, one HBM write, one memory round trip.
template <typename T, int ACT, bool HAL, int PPT, int NT>
__global__ void __launch_bounds__(NT, 512 / NT) norm_bwd_reg_kernel(dtg_plane dy, dtg_plane dy2, dtg_plane yp, dtg_plane x,
                                                              const float* __restrict__ stats,
                                                              const float* __restrict__ gamma, float* __restrict__ sums,
                                                              dtg_plane dx, dtg_plane dres, int mode, int nv, int chunk,
                                                              int dbg) {
  pdl_enter();
  constexpr int V = Vec<T>::N;
  __shared__ float red[NT * V * 2];
  __shared__ float allpart[kMaxCluster][kSlabCh * 2];
  __shared__ float4 kco[kSlabCh];
  cg::cluster_group cl = cg::this_cluster();
  const int cs = cl.num_blocks(), rank = cl.block_rank();
  const int slab_ch = nv * V;
  const int v = threadIdx.x % nv, lane = threadIdx.x / nv, lanes = NT / nv;
  const int n = blockIdx.y;
  const int c = blockIdx.x * slab_ch + v * V;
  const int hw = x.h * x.w;
  const int p0 = rank * chunk, p1 = min(hw, p0 + chunk);
  const size_t es = sizeof(T);
  const size_t pitch = static_cast<size_t>(x.c) * es;
  const uint8_t* xb = reinterpret_cast<const uint8_t*>(x.ptr) + (static_cast<size_t>(n) * hw * x.c + c) * es;
  const int hdy = dy.halo, hy = yp.ptr ? yp.halo : 0;
  const uint8_t* dyb = reinterpret_cast<const uint8_t*>(dy.ptr) +
                       (static_cast<size_t>(n) * (dy.h + 2 * hdy) * (dy.w + 2 * hdy) * dy.c + c) * es;
  const uint8_t* yb = yp.ptr ? reinterpret_cast<const uint8_t*>(yp.ptr) +
                                   (static_cast<size_t>(n) * (yp.h + 2 * hy) * (yp.w + 2 * hy) * yp.c + c) * es
                             : nullptr;
  const uint8_t* dy2b = dy2.ptr ? reinterpret_cast<const uint8_t*>(dy2.ptr) + (static_cast<size_t>(n) * hw * dy2.c + c) * es
                                : nullptr;
  // dx may carry a (zero) halo ring that this kernel never writes: the flat-raster dgrad of conv_patch2.cu reads it as padding
  const int hdx = dx.halo, wdx = dx.w + 2 * hdx;
  uint8_t* dxb = reinterpret_cast<uint8_t*>(dx.ptr) +
                 ((static_cast<size_t>(n) * (dx.h + 2 * hdx) * wdx + static_cast<size_t>(hdx) * wdx + hdx) * dx.c + c) * es;
  uint8_t* drb = dres.ptr ? reinterpret_cast<uint8_t*>(dres.ptr) + (static_cast<size_t>(n) * hw * dres.c + c) * es : nullptr;
  if (dbg & 2) drb = nullptr;

  // mode NONE = activation-only layer (conv bias + ReLU, modules.py:211-213): dx = g, sums = (sum g, -)
  const bool has_norm = mode != DTG_NORM_NONE;
  float mean[V], rstd[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    mean[i] = 0.f;
    rstd[i] = 0.f;
    if (has_norm) {
      const float2 mr = *reinterpret_cast<const float2*>(stats + (static_cast<size_t>(n) * x.c + c + i) * 2);
      mean[i] = mr.x;
      rstd[i] = mr.y;
    }
  }
  float g[PPT][V], xh[PPT][V];
  // phase A: every load of this thread is issued before any arithmetic (a warp issues in order: a use-after-load
  // in pixel j would otherwise serialise pixel j+1's loads behind a full memory round trip).  Out-of-range pixels
  // are clamped to the last valid one and zeroed afterwards; the rare reflect-fold extras come in phase A'.
  uint4 r_dy[PPT], r_d2[PPT], r_y[PPT], r_x[PPT];
  int bo0[PPT], bdr[PPT], bdc[PPT];
  {
    const int pl = p1 - 1;
    const int wdy = dy.w + 2 * hdy, wy = x.w + 2 * hy;
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      const int p = min(p0 + lane + j * lanes, pl);
      int py = 0, px = 0;
      if (HAL) {
        py = p / x.w;
        px = p - py * x.w;
      }
      const int o0 = HAL ? (py + hdy) * wdy + px + hdy : p;
      const int o1 = HAL ? (py + hy) * wy + px + hy : p;
      r_dy[j] = *reinterpret_cast<const uint4*>(dyb + static_cast<size_t>(o0) * pitch);
      if (dy2b) r_d2[j] = *reinterpret_cast<const uint4*>(dy2b + static_cast<size_t>(p) * pitch);
      if (ACT != DTG_ACT_NONE) r_y[j] = *reinterpret_cast<const uint4*>(yb + static_cast<size_t>(o1) * pitch);
      r_x[j] = has_norm ? *reinterpret_cast<const uint4*>(xb + static_cast<size_t>(p) * pitch) : make_uint4(0u, 0u, 0u, 0u);
      bo0[j] = o0;
      bdr[j] = (HAL && hdy == 1) ? (py == 1 ? -2 * wdy : (py == dy.h - 2 ? 2 * wdy : 0)) : 0;
      bdc[j] = (HAL && hdy == 1) ? (px == 1 ? -2 : (px == dy.w - 2 ? 2 : 0)) : 0;
    }
  }
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    Vec<T>::unpack(r_dy[j], g[j]);
    if (HAL && hdy == 1) {     // reflection-pad(1) backward (launcher guarantees h, w >= 4 and halo <= 1)
      if (bdr[j] != 0) {
        float t[V];
        Vec<T>::load(dyb + static_cast<size_t>(bo0[j] + bdr[j]) * pitch, t);
#pragma unroll
        for (int i = 0; i < V; ++i) g[j][i] += t[i];
      }
      if (bdc[j] != 0) {
        float t[V];
        Vec<T>::load(dyb + static_cast<size_t>(bo0[j] + bdc[j]) * pitch, t);
#pragma unroll
        for (int i = 0; i < V; ++i) g[j][i] += t[i];
        if (bdr[j] != 0) {
          Vec<T>::load(dyb + static_cast<size_t>(bo0[j] + bdr[j] + bdc[j]) * pitch, t);
#pragma unroll
          for (int i = 0; i < V; ++i) g[j][i] += t[i];
        }
      }
    }
    if (dy2b) {
      float t[V];
      Vec<T>::unpack(r_d2[j], t);
#pragma unroll
      for (int i = 0; i < V; ++i) g[j][i] += t[i];
    }
    if (ACT != DTG_ACT_NONE) {
      float t[V];
      Vec<T>::unpack(r_y[j], t);
#pragma unroll
      for (int i = 0; i < V; ++i) g[j][i] = t[i] > 0.f ? g[j][i] : (ACT == DTG_ACT_LRELU ? 0.2f * g[j][i] : 0.f);
    }
    Vec<T>::unpack(r_x[j], xh[j]);
    if (p0 + lane + j * lanes >= p1) {
#pragma unroll
      for (int i = 0; i < V; ++i) g[j][i] = 0.f;
    }
  }
  float s1[V], s2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) s1[i] = s2[i] = 0.f;
#pragma unroll
  for (int j = 0; j < PPT; ++j)
#pragma unroll
    for (int i = 0; i < V; ++i) {
      xh[j][i] = (xh[j][i] - mean[i]) * rstd[i];
      s1[i] += g[j][i];
      s2[i] += g[j][i] * xh[j][i];
    }
  block_reduce_push<V, NT>(s1, s2, nv, red, allpart, cl, cs, rank);
  cl.sync();
  if (threadIdx.x < slab_ch) {
    const int t = threadIdx.x;
    const float2 tot = cluster_total(allpart, cs, t);
    const float A = tot.x, B = tot.y;
    const int ch = blockIdx.x * slab_ch + t;
    const size_t nc = static_cast<size_t>(n) * x.c + ch;
    const float m = static_cast<float>(hw);
    const float d = mode == DTG_NORM_COND_INSTANCE ? m - 1.f : m;
    if (has_norm) {
      const float ga = mode == DTG_NORM_COND_INSTANCE ? gamma[nc] : gamma[ch];
      kco[t] = make_float4(stats[nc * 2 + 1] * ga, A / m, B / d, 0.f);
    } else {
      kco[t] = make_float4(1.f, 0.f, 0.f, 0.f);
    }
    if (rank == 0) {
      sums[nc * 2] = A;
      sums[nc * 2 + 1] = B;
    }
  }
  __syncthreads();
  float k0[V], kA[V], kB[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float4 k = kco[v * V + i];
    k0[i] = k.x;
    kA[i] = k.y;
    kB[i] = k.z;
  }
  int p = p0 + lane;
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    if (p < p1) {
      if (drb) Vec<T>::store(drb + p * pitch, g[j]);
      float o[V];
#pragma unroll
      for (int i = 0; i < V; ++i) o[i] = k0[i] * (g[j][i] - kA[i] - xh[j][i] * kB[i]);
      const size_t po = hdx ? static_cast<size_t>(p / x.w) * wdx + (p % x.w) : static_cast<size_t>(p);
      if (!(dbg & 4)) Vec<T>::store(dxb + po * pitch, o);
    }
    p += lanes;
  }
}

static const dtg_plane kNull = {nullptr, 0, 0, 0, 0, 0, 0};

constexpr int kRegPPT = 4;     // pixels per thread of the register-resident backward
constexpr int kRegNT = 256;    // its threads per CTA (512 x 2 pixels measured slower: 83 vs 69 us)

struct FusedGeom {
  int nv, cblocks, cs, chunk;
  int reg_cs, reg_chunk;       // register-resident backward: cluster size / pixels per CTA (0 = not applicable)
  int reg_nt, reg_nv, reg_cblocks;
};

static bool fused_geom(const dtg_plane* x, FusedGeom* g) {
  static const bool disabled = getenv("DTG_NO_FUSED_NORM") != nullptr;
  if (disabled) return false;
  const int es = elem_size(x->dtype);
  const int rowb = x->c * es;
  static const int slab_env = getenv("DTG_FUSED_SLAB") ? atoi(getenv("DTG_FUSED_SLAB")) : 128;
  const int slabb = std::min(rowb, slab_env);
  if (slabb != 32 && slabb != 64 && slabb != 128) return false;
  if (rowb % slabb != 0) return false;
  const int hw = x->h * x->w;
  if (hw < 2) return false;
  g->nv = slabb / 16;
  g->cblocks = rowb / slabb;
  int cs = 1;
  while (cs < 8 && hw / (cs * 2) >= 384) cs *= 2;
  g->cs = cs;
  g->chunk = (hw + cs - 1) / cs;
  // register-resident backward: one CTA covers lanes * kRegPPT pixels, a cluster (<= 8 CTAs) covers the slab
  // 128-thread CTAs on 64-byte slabs (4 clusters in flight per SM instead of 2, so that one cluster's load phase
  // overlaps another's reduce / store phase): 50 vs 60 us on [80,32,32,128] + residual.  DTG_REG_VAR: 0 = 256 threads
  // on 128-byte slabs, 2 = 64 threads on 32-byte slabs.
  static const int reg_var = getenv("DTG_REG_VAR") ? atoi(getenv("DTG_REG_VAR")) : 1;
  g->reg_nt = kRegNT;
  g->reg_nv = g->nv;
  g->reg_cblocks = g->cblocks;
  if (reg_var == 1 && slabb >= 64) {
    g->reg_nt = 128;
    g->reg_nv = 4;
    g->reg_cblocks = rowb / 64;
  } else if (reg_var == 2) {
    g->reg_nt = 64;
    g->reg_nv = 2;
    g->reg_cblocks = rowb / 32;
  }
  const int per_cta = (g->reg_nt / g->reg_nv) * kRegPPT;
  g->reg_cs = 0;
  for (int r = 1; r <= 8; r *= 2)
    if (r * per_cta >= hw) {
      g->reg_cs = r;
      g->reg_chunk = (hw + r - 1) / r;
      break;
    }
  return x->n <= 65535;
}

template <typename K, typename... Args>
static int launch_cluster_nt(K kernel, dim3 grid, int nt, int cs, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = dim3(nt, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 1;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = cs;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  static const bool dbg_occ = getenv("DTG_DEBUG_OCC") != nullptr;
  if (dbg_occ) {
    int ncl = -1;
    cudaOccupancyMaxActiveClusters(&ncl, kernel, &cfg);
    fprintf(stderr, "[dtg] cluster launch grid (%u,%u,%u) cs %d: max active clusters %d (= %d CTAs)\n", grid.x, grid.y, grid.z,
            cs, ncl, ncl * cs);
  }
  DTG_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, args...));
  DTG_LAUNCHED();
  return DTG_OK;
}

template <typename K, typename... Args>
static int launch_cluster(K kernel, dim3 grid, int cs, cudaStream_t stream, Args... args) {
  return launch_cluster_nt(kernel, grid, kFT, cs, stream, args...);
}

int try_norm_fwd_fused(const dtg_norm_args* a, const dtg_plane* x, const dtg_plane* residual, const float* gamma,
                       const float* beta, float* stats, const dtg_plane* out, cudaStream_t stream) {
  if (a->phase != 0 || (a->mode != DTG_NORM_INSTANCE && a->mode != DTG_NORM_COND_INSTANCE)) return 1;
  FusedGeom g;
  if (!fused_geom(x, &g)) return 1;
  const dtg_plane res = (residual && residual->ptr) ? *residual : kNull;
  const dim3 grid(g.cblocks, x->n, g.cs);
  if (x->dtype == DTG_BF16)
    return launch_cluster(norm_fwd_fused_kernel<__nv_bfloat16>, grid, g.cs, stream, *x, res, *out, gamma, beta, stats,
                          static_cast<int>(a->mode), static_cast<int>(a->act), a->eps, g.nv, g.chunk);
  return launch_cluster(norm_fwd_fused_kernel<float>, grid, g.cs, stream, *x, res, *out, gamma, beta, stats,
                        static_cast<int>(a->mode), static_cast<int>(a->act), a->eps, g.nv, g.chunk);
}

int try_norm_bwd_fused(const dtg_norm_args* a, const dtg_plane* dy, const dtg_plane* dy2, const dtg_plane* y,
                       const dtg_plane* x, const float* stats, const float* gamma, float* sums, const dtg_plane* dx,
                       const dtg_plane* d_res, cudaStream_t stream) {
  if (a->phase != 0 || a->mode == DTG_NORM_BATCH) return 1;
  FusedGeom g;
  if (!fused_geom(x, &g)) return 1;
  if (a->mode == DTG_NORM_NONE && g.reg_cs == 0) return 1;     // activation-only: register-resident kernel only
  const dtg_plane p_dy2 = (dy2 && dy2->ptr) ? *dy2 : kNull;
  const dtg_plane p_y = (y && y->ptr) ? *y : kNull;
  const dtg_plane p_res = (d_res && d_res->ptr) ? *d_res : kNull;
  const dim3 grid(g.cblocks, x->n, g.cs);
  DTG_REQUIRE(dy->c == x->c && dx->c == x->c && (!p_y.ptr || p_y.c == x->c) && (!p_dy2.ptr || p_dy2.c == x->c) &&
                  (!p_res.ptr || p_res.c == x->c),
              "norm_bwd_fused: channel mismatch");
  const bool hal = dy->halo > 0 || (p_y.ptr && p_y.halo > 0);
  static const bool no_reg = getenv("DTG_NO_REG_NORM") != nullptr;
  static const int dbgv = getenv("DTG_NORM_DBG") ? atoi(getenv("DTG_NORM_DBG")) : 0;
  const dim3 rgrid(g.reg_cblocks, x->n, g.reg_cs);
  const bool reg_ok = dy->halo == 0 || (dy->halo == 1 && dy->h >= 4 && dy->w >= 4);
  if (dx->halo > 0 && !(g.reg_cs > 0 && !no_reg && reg_ok)) {
    set_error("norm_bwd: a dx plane with a halo is only supported by the register-resident kernel (<= %d pixels per slab)",
              8 * (g.reg_nt / g.reg_nv) * kRegPPT);
    return DTG_ERR_INVALID;
  }
#define DTG_REG_LAUNCH(TT, AA, NTV)                                                                                     \
  do {                                                                                                                  \
    if (hal)                                                                                                            \
      return launch_cluster_nt(norm_bwd_reg_kernel<TT, AA, true, kRegPPT, NTV>, rgrid, NTV, g.reg_cs, stream, *dy,      \
                               p_dy2, p_y, *x, stats, gamma, sums, *dx, p_res, static_cast<int>(a->mode), g.reg_nv,     \
                               g.reg_chunk, dbgv);                                                                      \
    return launch_cluster_nt(norm_bwd_reg_kernel<TT, AA, false, kRegPPT, NTV>, rgrid, NTV, g.reg_cs, stream, *dy,       \
                             p_dy2, p_y, *x, stats, gamma, sums, *dx, p_res, static_cast<int>(a->mode), g.reg_nv,       \
                             g.reg_chunk, dbgv);                                                                        \
  } while (0)
#define DTG_BWD_LAUNCH(TT, AA)                                                                                          \
  do {                                                                                                                  \
    if (a->mode == DTG_NORM_NONE && !(g.reg_cs > 0 && !no_reg && reg_ok)) return 1;                                     \
    if (g.reg_cs > 0 && !no_reg && reg_ok) {                                                                            \
      if (g.reg_nt == 128) DTG_REG_LAUNCH(TT, AA, 128);                                                                 \
      if (g.reg_nt == 64) DTG_REG_LAUNCH(TT, AA, 64);                                                                   \
      DTG_REG_LAUNCH(TT, AA, 256);                                                                                      \
    }                                                                                                                   \
    if (hal)                                                                                                            \
      return launch_cluster(norm_bwd_fused_kernel<TT, AA, true>, grid, g.cs, stream, *dy, p_dy2, p_y, *x, stats, gamma,    \
                            sums, *dx, p_res, static_cast<int>(a->mode), g.nv, g.chunk);                                 \
    return launch_cluster(norm_bwd_fused_kernel<TT, AA, false>, grid, g.cs, stream, *dy, p_dy2, p_y, *x, stats, gamma,     \
                          sums, *dx, p_res, static_cast<int>(a->mode), g.nv, g.chunk);                                   \
  } while (0)
  if (x->dtype == DTG_BF16) {
    if (a->act == DTG_ACT_RELU) DTG_BWD_LAUNCH(__nv_bfloat16, DTG_ACT_RELU);
    if (a->act == DTG_ACT_LRELU) DTG_BWD_LAUNCH(__nv_bfloat16, DTG_ACT_LRELU);
    if (a->act == DTG_ACT_NONE) DTG_BWD_LAUNCH(__nv_bfloat16, DTG_ACT_NONE);
  } else {
    if (a->act == DTG_ACT_RELU) DTG_BWD_LAUNCH(float, DTG_ACT_RELU);
    if (a->act == DTG_ACT_LRELU) DTG_BWD_LAUNCH(float, DTG_ACT_LRELU);
    if (a->act == DTG_ACT_NONE) DTG_BWD_LAUNCH(float, DTG_ACT_NONE);
  }
#undef DTG_BWD_LAUNCH
#undef DTG_REG_LAUNCH
  return 1;
}

}  // namespace dtg
