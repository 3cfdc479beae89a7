// Fused (conditional) instance / batch normalisation + affine + activation (+ residual, + reflect
// halo) forward and backward on NHWC planes.  HBM-bound: every pass streams 16-byte vectors along the
// contiguous channel axis; statistics are reduced with a fixed-order block reduction (deterministic,
// no atomics).  Formulas: SURVEY.md 9.1 (verified against autograd in fp64).
//
// forward : stats pass (read x) -> finalize ([n][c] coefficients) -> apply pass (read x, write y)
// backward: reduce pass (read dy,y,x) -> finalize -> apply pass (read dy,y,x, write dx [, d_res])
#include <algorithm>

#include "common.cuh"
#include "norm_common.cuh"

namespace dtg {

constexpr int kNormThreads = 256;
constexpr int kMaxSplits = 32;

// Block-level fixed-order reduction of per-thread (s1[V], s2[V]) over the pixel lanes.
// Thread t owns vector column v = t % nv; result for (v, i) written by thread (v*V + i) < nv*V.
template <int V>
__device__ __forceinline__ void block_reduce_store(float (&s1)[V], float (&s2)[V], int nv, float* smem, float* out_c0,
                                                   int cbase) {
  const int t = threadIdx.x;
  const int lanes = kNormThreads / nv;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    smem[(t * V + i) * 2] = s1[i];
    smem[(t * V + i) * 2 + 1] = s2[i];
  }
  __syncthreads();
  if (t < nv * V) {
    const int v = t / V, i = t % V;
    float a = 0.f, b = 0.f;
    for (int l = 0; l < lanes; ++l) {
      const int src = l * nv + v;
      a += smem[(src * V + i) * 2];
      b += smem[(src * V + i) * 2 + 1];
    }
    out_c0[(cbase + t) * 2] = a;
    out_c0[(cbase + t) * 2 + 1] = b;
  }
}

// forward statistics: partial[s][n][c] = (sum(x-K), sum((x-K)^2)), K = x[n,0,0,c] (0 for batch norm)
template <typename T>
__global__ void __launch_bounds__(kNormThreads) norm_stats_kernel(dtg_plane x, int cg, int splits, int use_shift,
                                                                  float* __restrict__ partial) {
  pdl_enter();
  constexpr int V = Vec<T>::N;
  __shared__ float smem[kNormThreads * V * 2];
  const int nv = cg / V;
  const int v = threadIdx.x % nv, lane = threadIdx.x / nv, lanes = kNormThreads / nv;
  const int n = blockIdx.y, s = blockIdx.z;
  const int c = blockIdx.x * cg + v * V;
  const int hw = x.h * x.w;
  const int chunk = (hw + splits - 1) / splits;
  const int p0 = s * chunk, p1 = min(hw, p0 + chunk);
  const uint8_t* base = reinterpret_cast<const uint8_t*>(x.ptr);
  float K[V], s1[V], s2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) K[i] = s1[i] = s2[i] = 0.f;
  if (use_shift) Vec<T>::load(base + (plane_pix(x, n, 0, 0) * x.c + c) * sizeof(T), K);
  for (int p = p0 + lane; p < p1; p += lanes) {
    float f[V];
    Vec<T>::load(base + (plane_pix(x, n, p / x.w, p % x.w) * x.c + c) * sizeof(T), f);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float d = f[i] - K[i];
      s1[i] += d;
      s2[i] += d * d;
    }
  }
  block_reduce_store<V>(s1, s2, nv, smem, partial + (static_cast<size_t>(s) * x.n + n) * x.c * 2, blockIdx.x * cg);
}

// backward sums: partial[s][n][c] = (sum g, sum g*xhat)
template <typename T>
__global__ void __launch_bounds__(kNormThreads) norm_bwd_reduce_kernel(dtg_plane dy, dtg_plane dy2, dtg_plane yp,
                                                                       dtg_plane x, const float* __restrict__ stats,
                                                                       int mode, int act, int cg, int splits,
                                                                       float* __restrict__ partial) {
  pdl_enter();
  constexpr int V = Vec<T>::N;
  __shared__ float smem[kNormThreads * V * 2];
  const int nv = cg / V;
  const int v = threadIdx.x % nv, lane = threadIdx.x / nv, lanes = kNormThreads / nv;
  const int n = blockIdx.y, s = blockIdx.z;
  const int c = blockIdx.x * cg + v * V;
  const int hw = x.h * x.w;
  const int chunk = (hw + splits - 1) / splits;
  const int p0 = s * chunk, p1 = min(hw, p0 + chunk);
  float mean[V], rstd[V], s1[V], s2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    s1[i] = s2[i] = 0.f;
    mean[i] = 0.f;
    rstd[i] = 0.f;
    if (mode != DTG_NORM_NONE) {
      mean[i] = stats[(static_cast<size_t>(n) * x.c + c + i) * 2];
      rstd[i] = stats[(static_cast<size_t>(n) * x.c + c + i) * 2 + 1];
    }
  }
  for (int p = p0 + lane; p < p1; p += lanes) {
    const int py = p / x.w, px = p % x.w;
    float g[V];
    load_g<T>(dy, dy2, yp, act, n, py, px, c, g);
    if (mode != DTG_NORM_NONE) {
      float f[V];
      Vec<T>::load(reinterpret_cast<const uint8_t*>(x.ptr) + (plane_pix(x, n, py, px) * x.c + c) * sizeof(T), f);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        s1[i] += g[i];
        s2[i] += g[i] * (f[i] - mean[i]) * rstd[i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) s1[i] += g[i];
    }
  }
  block_reduce_store<V>(s1, s2, nv, smem, partial + (static_cast<size_t>(s) * x.n + n) * x.c * 2, blockIdx.x * cg);
}

// batch norm: collapse partial[s][n][c] over (s, n) -> bnsum[c][2]
__global__ void bn_collapse_kernel(const float* __restrict__ partial, int splits, int n, int c, float* __restrict__ bnsum) {
  pdl_enter();
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  float a = 0.f, b = 0.f;
  for (int s = 0; s < splits; ++s)
    for (int i = 0; i < n; ++i) {
      const size_t o = ((static_cast<size_t>(s) * n + i) * c + ch) * 2;
      a += partial[o];
      b += partial[o + 1];
    }
  bnsum[ch * 2] = a;
  bnsum[ch * 2 + 1] = b;
}

// forward finalize: stats[n][c] = (mean, rstd); coef[n][c] = (a, b) with y = x*a + b
__global__ void norm_fwd_finalize_kernel(const float* __restrict__ partial, const float* __restrict__ bnsum,
                                         const void* xptr, int dtype, int splits, int n, int c, int hw, int mode,
                                         float eps, float momentum, int world, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, float* __restrict__ bn_running,
                                         float* __restrict__ stats, float* __restrict__ coef) {
  pdl_enter();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (mode == DTG_NORM_BATCH) {
    if (idx >= c) return;
    const float m = static_cast<float>(hw) * n * world;
    const float mean = bnsum[idx * 2] / m;
    float var = bnsum[idx * 2 + 1] / m - mean * mean;
    var = var < 0.f ? 0.f : var;
    const float rstd = rsqrtf(var + eps);
    if (bn_running) {
      const float unb = m > 1.f ? var * m / (m - 1.f) : var;
      bn_running[idx] = (1.f - momentum) * bn_running[idx] + momentum * mean;
      bn_running[c + idx] = (1.f - momentum) * bn_running[c + idx] + momentum * unb;
    }
    const float a = rstd * gamma[idx], b = beta[idx] - mean * a;
    for (int i = 0; i < n; ++i) {
      stats[(static_cast<size_t>(i) * c + idx) * 2] = mean;
      stats[(static_cast<size_t>(i) * c + idx) * 2 + 1] = rstd;
      coef[(static_cast<size_t>(i) * c + idx) * 2] = a;
      coef[(static_cast<size_t>(i) * c + idx) * 2 + 1] = b;
    }
    return;
  }
  if (idx >= n * c) return;
  const int ch = idx % c, i = idx / c;
  float s1 = 0.f, s2 = 0.f;
  for (int s = 0; s < splits; ++s) {
    const size_t o = ((static_cast<size_t>(s) * n + i) * c + ch) * 2;
    s1 += partial[o];
    s2 += partial[o + 1];
  }
  const size_t first = static_cast<size_t>(i) * hw * c + ch;  // x has halo 0
  const float K = dtype == DTG_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(xptr)[first])
                                    : reinterpret_cast<const float*>(xptr)[first];
  const float m = static_cast<float>(hw);
  const float mean = K + s1 / m;
  const float d = mode == DTG_NORM_COND_INSTANCE ? m - 1.f : m;
  float var = (s2 - s1 * s1 / m) / d;
  var = var < 0.f ? 0.f : var;
  const float rstd = rsqrtf(var + eps);
  stats[idx * 2] = mean;
  stats[idx * 2 + 1] = rstd;
  const float ga = mode == DTG_NORM_COND_INSTANCE ? gamma[idx] : gamma[ch];
  const float be = mode == DTG_NORM_COND_INSTANCE ? beta[idx] : beta[ch];
  const float a = rstd * ga;
  coef[idx * 2] = a;
  coef[idx * 2 + 1] = be - mean * a;
}

// forward apply: y = act(x*a + b (+ residual)), mirrored into the output halo
template <typename T>
__global__ void __launch_bounds__(kNormThreads) norm_apply_kernel(dtg_plane x, dtg_plane res, const float* __restrict__ coef,
                                                                  int has_coef, int act, dtg_plane out) {
  pdl_enter();
  constexpr int V = Vec<T>::N;
  const int nvc = x.c / V;
  const size_t total = static_cast<size_t>(x.n) * x.h * x.w * nvc;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % nvc) * V;
    size_t r = idx / nvc;
    const int px = r % x.w;
    r /= x.w;
    const int py = r % x.h;
    const int n = r / x.h;
    float f[V];
    Vec<T>::load(reinterpret_cast<const uint8_t*>(x.ptr) + (plane_pix(x, n, py, px) * x.c + c) * sizeof(T), f);
    if (has_coef) {
      const float2* cf = reinterpret_cast<const float2*>(coef) + static_cast<size_t>(n) * x.c + c;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float2 ab = __ldg(cf + i);
        f[i] = f[i] * ab.x + ab.y;
      }
    }
    if (res.ptr) {
      float t[V];
      Vec<T>::load(reinterpret_cast<const uint8_t*>(res.ptr) + (plane_pix(res, n, py, px) * res.c + c) * sizeof(T), t);
#pragma unroll
      for (int i = 0; i < V; ++i) f[i] += t[i];
    }
#pragma unroll
    for (int i = 0; i < V; ++i) f[i] = apply_act(f[i], act);
    int hts[3], wts[3];
    const int nh = reflect_targets(py, out.h, out.halo, hts), nw = reflect_targets(px, out.w, out.halo, wts);
    for (int a = 0; a < nh; ++a)
      for (int q = 0; q < nw; ++q)
        Vec<T>::store(reinterpret_cast<uint8_t*>(out.ptr) + (plane_pix(out, n, hts[a], wts[q]) * out.c + c) * sizeof(T), f);
  }
}

// backward finalize, stage 1 (one thread per (n,c)): sums[n][c] = (A, B) summed over the pixel splits in a
// fixed order; instance modes also emit kcoef[n][c] = (rstd*gamma, A/m, B/d, 0).
__global__ void norm_bwd_sums_kernel(const float* __restrict__ partial, int splits, int n, int c, int hw, int mode,
                                     const float* __restrict__ stats, const float* __restrict__ gamma,
                                     float* __restrict__ sums, float* __restrict__ kcoef) {
  pdl_enter();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * c) return;
  const int ch = idx % c;
  float A = 0.f, B = 0.f;
  for (int s = 0; s < splits; ++s) {
    const size_t o = (static_cast<size_t>(s) * n * c + idx) * 2;
    A += partial[o];
    B += partial[o + 1];
  }
  sums[idx * 2] = A;
  sums[idx * 2 + 1] = B;
  if (mode == DTG_NORM_INSTANCE || mode == DTG_NORM_COND_INSTANCE) {
    const float m = static_cast<float>(hw);
    const float d = mode == DTG_NORM_COND_INSTANCE ? m - 1.f : m;
    const float ga = mode == DTG_NORM_COND_INSTANCE ? gamma[idx] : gamma[ch];
    kcoef[idx * 4] = stats[idx * 2 + 1] * ga;
    kcoef[idx * 4 + 1] = A / m;
    kcoef[idx * 4 + 2] = B / d;
  }
}

// stage 2: reduce sums over n -> parameter gradients (+=) and, for batch norm, the per-channel sums handed to the
// cross-GPU all-reduce.  Block = 32 channels x 8 sample lanes (fixed-order two-level sum: deterministic).
__global__ void __launch_bounds__(256) norm_bwd_channel_kernel(const float* __restrict__ sums, int n, int c, int mode,
                                                                 float* __restrict__ bnsum, float* __restrict__ d_gamma,
                                                                 float* __restrict__ d_beta) {
  pdl_enter();
  __shared__ float2 sh[8][32];
  const int cl = threadIdx.x & 31, nl = threadIdx.x >> 5;
  const int ch = blockIdx.x * 32 + cl;
  float a = 0.f, b = 0.f;
  if (ch < c) {
#pragma unroll 4
    for (int i = nl; i < n; i += 8) {
      const float2 v = *reinterpret_cast<const float2*>(sums + (static_cast<size_t>(i) * c + ch) * 2);
      a += v.x;
      b += v.y;
    }
  }
  sh[nl][cl] = make_float2(a, b);
  __syncthreads();
  if (nl != 0 || ch >= c) return;
  a = b = 0.f;
#pragma unroll
  for (int l = 0; l < 8; ++l) {
    a += sh[l][cl].x;
    b += sh[l][cl].y;
  }
  if (mode == DTG_NORM_BATCH) {
    bnsum[ch * 2] = a;
    bnsum[ch * 2 + 1] = b;
  }
  if (mode != DTG_NORM_COND_INSTANCE) {   // parameter gradients use the LOCAL sums (the caller all-reduces them)
    if (d_beta) d_beta[ch] += a;
    if (d_gamma && mode != DTG_NORM_NONE) d_gamma[ch] += b;
  }
}

// stage 3, batch norm only (one thread per (n,c)): kcoef from the (all-reduced) per-channel sums
__global__ void norm_bwd_bn_kcoef_kernel(const float* __restrict__ bnsum, int n, int c, int hw, int world,
                                         const float* __restrict__ stats, const float* __restrict__ gamma,
                                         float* __restrict__ kcoef) {
  pdl_enter();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * c) return;
  const int ch = idx % c;
  const float m = static_cast<float>(hw) * n * world;
  kcoef[idx * 4] = stats[idx * 2 + 1] * gamma[ch];
  kcoef[idx * 4 + 1] = bnsum[ch * 2] / m;
  kcoef[idx * 4 + 2] = bnsum[ch * 2 + 1] / m;
}

// backward apply: dx = k0*(g - kA - xhat*kB)  (NONE: dx = g); d_res = g
template <typename T>
__global__ void __launch_bounds__(kNormThreads) norm_bwd_apply_kernel(dtg_plane dy, dtg_plane dy2, dtg_plane yp, dtg_plane x,
                                                                      const float* __restrict__ stats,
                                                                      const float* __restrict__ kcoef, int mode, int act,
                                                                      dtg_plane dx, dtg_plane dres) {
  pdl_enter();
  constexpr int V = Vec<T>::N;
  const int nvc = dx.c / V;
  const size_t total = static_cast<size_t>(dx.n) * dx.h * dx.w * nvc;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % nvc) * V;
    size_t r = idx / nvc;
    const int px = r % dx.w;
    r /= dx.w;
    const int py = r % dx.h;
    const int n = r / dx.h;
    float g[V];
    load_g<T>(dy, dy2, yp, act, n, py, px, c, g);
    if (dres.ptr)
      Vec<T>::store(reinterpret_cast<uint8_t*>(dres.ptr) + (plane_pix(dres, n, py, px) * dres.c + c) * sizeof(T), g);
    if (mode != DTG_NORM_NONE) {
      float f[V];
      Vec<T>::load(reinterpret_cast<const uint8_t*>(x.ptr) + (plane_pix(x, n, py, px) * x.c + c) * sizeof(T), f);
      const float2* st = reinterpret_cast<const float2*>(stats) + static_cast<size_t>(n) * x.c + c;
      const float4* kc = reinterpret_cast<const float4*>(kcoef) + static_cast<size_t>(n) * x.c + c;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float2 mr = __ldg(st + i);
        const float4 k = __ldg(kc + i);
        const float xh = (f[i] - mr.x) * mr.y;
        g[i] = k.x * (g[i] - k.y - xh * k.z);
      }
    }
    Vec<T>::store(reinterpret_cast<uint8_t*>(dx.ptr) + (plane_pix(dx, n, py, px) * dx.c + c) * sizeof(T), g);
  }
}

// ---- CondInstanceNorm z projections ----------------------------------------------------------
__global__ void cin_affine_fwd_kernel(const float* __restrict__ z, const float* __restrict__ ws, const float* __restrict__ bs,
                                      const float* __restrict__ wb, const float* __restrict__ bb, int n, int c, int nz,
                                      float* __restrict__ gamma, float* __restrict__ beta) {
  pdl_enter();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * c) return;
  const int ch = idx % c, i = idx / c;
  float s = bs[ch], b = bb[ch];
  for (int k = 0; k < nz; ++k) {
    const float zz = z[i * nz + k];
    s += ws[ch * nz + k] * zz;
    b += wb[ch * nz + k] * zz;
  }
  gamma[idx] = s > 0.f ? s : 0.f;
  beta[idx] = b > 0.f ? b : 0.f;
}

// Blocks [0, c): one channel each -- 16 latent columns x 8 sample lanes reduce over n (fixed-order two-level sum).
// Blocks [c, c+n): one sample each -- 16 latent columns x 8 channel lanes reduce over c for dz.  nz <= 16.
__global__ void __launch_bounds__(128) cin_affine_bwd_kernel(const float* __restrict__ z, const float* __restrict__ ws,
                                                               const float* __restrict__ wb, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, const float* __restrict__ sums,
                                                               int n, int c, int nz, float* __restrict__ d_ws,
                                                               float* __restrict__ d_bs, float* __restrict__ d_wb,
                                                               float* __restrict__ d_bb, float* __restrict__ d_z) {
  pdl_enter();
  __shared__ float4 sh[8][16];
  const int l = threadIdx.x >> 4;
  for (int k0 = 0; k0 < nz; k0 += 16) {     // latent columns in groups of 16
    const int k = k0 + (threadIdx.x & 15), kl = threadIdx.x & 15;
    if (blockIdx.x < c) {
      const int ch = blockIdx.x;
      float aw = 0.f, ab = 0.f, sw = 0.f, sb = 0.f;
      for (int i = l; i < n; i += 8) {
        const size_t o = static_cast<size_t>(i) * c + ch;
        const float2 su = *reinterpret_cast<const float2*>(sums + o * 2);
        const float ds = gamma[o] > 0.f ? su.y : 0.f;   // d_scale = sum g*xhat
        const float db = beta[o] > 0.f ? su.x : 0.f;    // d_shift = sum g
        const float zz = k < nz ? z[i * nz + k] : 0.f;
        aw += ds * zz;
        ab += db * zz;
        sw += ds;
        sb += db;
      }
      sh[l][kl] = make_float4(aw, ab, sw, sb);
      __syncthreads();
      if (l == 0 && k < nz) {
        float4 t = sh[0][kl];
        for (int j = 1; j < 8; ++j) {
          t.x += sh[j][kl].x;
          t.y += sh[j][kl].y;
          t.z += sh[j][kl].z;
          t.w += sh[j][kl].w;
        }
        d_ws[ch * nz + k] += t.x;
        d_wb[ch * nz + k] += t.y;
        if (k == 0) {
          d_bs[ch] += t.z;
          d_bb[ch] += t.w;
        }
      }
    } else if (d_z != nullptr) {
      const int i = blockIdx.x - c;
      float acc = 0.f;
      for (int ch = l; ch < c; ch += 8) {
        const size_t o = static_cast<size_t>(i) * c + ch;
        const float2 su = *reinterpret_cast<const float2*>(sums + o * 2);
        const float ds = gamma[o] > 0.f ? su.y : 0.f;
        const float db = beta[o] > 0.f ? su.x : 0.f;
        if (k < nz) acc += ds * ws[ch * nz + k] + db * wb[ch * nz + k];
      }
      sh[l][kl].x = acc;
      __syncthreads();
      if (l == 0 && k < nz) {
        float t = sh[0][kl].x;
        for (int j = 1; j < 8; ++j) t += sh[j][kl].x;
        d_z[i * nz + k] += t;
      }
    }
    __syncthreads();
  }
}

static int pick_splits(int n, int c, int cg, int hw) {
  const int ctas = n * (c / cg);
  int s = (6 * 148 + ctas - 1) / ctas;
  s = std::min(s, kMaxSplits);
  s = std::min(s, std::max(1, hw / 64));
  return std::max(1, s);
}

static int elemwise_grid(size_t total) {
  size_t g = (total + kNormThreads - 1) / kNormThreads;
  return static_cast<int>(std::max<size_t>(1, std::min<size_t>(g, 148 * 32)));
}

static const dtg_plane kNullPlane = {nullptr, 0, 0, 0, 0, 0, 0};

}  // namespace dtg

using namespace dtg;

extern "C" size_t dtg_norm_workspace_bytes(const dtg_plane* x) {
  if (!x) return 0;
  const size_t nc = static_cast<size_t>(x->n) * x->c;
  return (2 * static_cast<size_t>(x->c) + kMaxSplits * nc * 2 + nc * 4 + nc * 2) * sizeof(float);
}

extern "C" int dtg_norm_fwd(const dtg_norm_args* a, const dtg_plane* x, const dtg_plane* residual, const float* gamma,
                            const float* beta, float* bn_running, float* stats, float* coef, float* partial,
                            const dtg_plane* out, void* stream_) {
  DTG_REQUIRE(a && x && out && x->ptr && out->ptr, "dtg_norm_fwd: null argument");
  DTG_REQUIRE(x->halo == 0, "dtg_norm_fwd: x must have halo 0");
  DTG_REQUIRE(x->dtype == out->dtype && x->n == out->n && x->h == out->h && x->w == out->w && x->c == out->c,
              "dtg_norm_fwd: out plane mismatch");
  DTG_REQUIRE(!residual || !residual->ptr || (residual->c == x->c && residual->h == x->h && residual->dtype == x->dtype),
              "dtg_norm_fwd: residual mismatch");
  const bool bf = x->dtype == DTG_BF16;
  const int V = bf ? 8 : 4;
  DTG_REQUIRE(x->c % V == 0, "dtg_norm_fwd: channels %d not a multiple of %d", x->c, V);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int hw = x->h * x->w;
  const int mode = a->mode;
  if (mode != DTG_NORM_NONE) {
    DTG_REQUIRE(gamma && beta && stats && coef && partial, "dtg_norm_fwd: missing buffers");
    DTG_REQUIRE(mode != DTG_NORM_COND_INSTANCE || hw > 1, "dtg_norm_fwd: conditional instance norm needs H*W > 1");
    {
      int rc = try_norm_fwd_lean(a, x, residual, gamma, beta, stats, partial + 2 * x->c, out, stream);
      if (rc == 1) rc = try_norm_fwd_tma(a, x, residual, gamma, beta, stats, out, stream);
      if (rc == 1) rc = try_norm_fwd_fused(a, x, residual, gamma, beta, stats, out, stream);
      if (rc <= 0) return rc;    // launched (0) or failed (<0); 1 = not handled by the cluster kernel
    }
    int cg = 8 * V;
    while (x->c % cg != 0) cg >>= 1;
    const int splits = pick_splits(x->n, x->c, cg, hw);
    float* bnsum = partial;
    float* part = partial + 2 * x->c;
    if (a->phase == 0 || a->phase == 1) {
      dim3 grid(x->c / cg, x->n, splits);
      if (bf)
        DTG_CHECK_CUDA(launch_k(norm_stats_kernel<__nv_bfloat16>, grid, kNormThreads, 0, stream, *x, cg, splits, mode != DTG_NORM_BATCH, part));
      else
        DTG_CHECK_CUDA(launch_k(norm_stats_kernel<float>, grid, kNormThreads, 0, stream, *x, cg, splits, mode != DTG_NORM_BATCH, part));
      if (mode == DTG_NORM_BATCH) {
        DTG_CHECK_CUDA(launch_k(bn_collapse_kernel, (x->c + 127) / 128, 128, 0, stream, part, splits, x->n, x->c, bnsum));
      }
      if (a->phase == 1) return DTG_OK;
    }
    const int world = a->world_size > 0 ? a->world_size : 1;
    const int cnt = mode == DTG_NORM_BATCH ? x->c : x->n * x->c;
    DTG_CHECK_CUDA(launch_k(norm_fwd_finalize_kernel, (cnt + 127) / 128, 128, 0, stream, part, bnsum, x->ptr, x->dtype, splits, x->n, x->c, hw,
                                                                   mode, a->eps, a->momentum, world, gamma, beta,
                                                                   bn_running, stats, coef));
  } else if (a->phase == 1) {
    return DTG_OK;
  }
  const dtg_plane res = (residual && residual->ptr) ? *residual : kNullPlane;
  const size_t total = static_cast<size_t>(x->n) * hw * (x->c / V);
  if (bf)
    DTG_CHECK_CUDA(launch_k(norm_apply_kernel<__nv_bfloat16>, elemwise_grid(total), kNormThreads, 0, stream, *x, res, coef, mode != DTG_NORM_NONE, a->act, *out));
  else
    DTG_CHECK_CUDA(launch_k(norm_apply_kernel<float>, elemwise_grid(total), kNormThreads, 0, stream, *x, res, coef, mode != DTG_NORM_NONE, a->act, *out));
  return DTG_OK;
}

extern "C" int dtg_norm_bwd(const dtg_norm_args* a, const dtg_plane* dy, const dtg_plane* dy2, const dtg_plane* y,
                            const dtg_plane* x, const float* stats, const float* gamma, float* sums, float* d_gamma,
                            float* d_beta, float* partial, const dtg_plane* dx, const dtg_plane* d_res, void* stream_) {
  DTG_REQUIRE(a && dy && dx && dy->ptr && dx->ptr && partial, "dtg_norm_bwd: null argument");
  const int mode = a->mode;
  DTG_REQUIRE(mode == DTG_NORM_NONE || (x && x->ptr && stats && gamma), "dtg_norm_bwd: missing saved tensors");
  DTG_REQUIRE(a->act == DTG_ACT_NONE || (y && y->ptr), "dtg_norm_bwd: activation needs saved output");
  DTG_REQUIRE(dx->halo >= 0 && dx->n == dy->n && dx->h == dy->h && dx->w == dy->w && dx->c == dy->c && dx->dtype == dy->dtype,
              "dtg_norm_bwd: dx / dy mismatch");
  DTG_REQUIRE(!x || !x->ptr || (x->halo == 0 && x->c == dx->c && x->h == dx->h), "dtg_norm_bwd: x mismatch");
  DTG_REQUIRE(!dy2 || !dy2->ptr || (dy2->c == dx->c && dy2->h == dx->h && dy2->halo == 0), "dtg_norm_bwd: dy2 mismatch");
  DTG_REQUIRE(!y || !y->ptr || (y->c == dx->c && y->h == dx->h), "dtg_norm_bwd: y mismatch");
  const bool bf = dx->dtype == DTG_BF16;
  const int V = bf ? 8 : 4;
  DTG_REQUIRE(dx->c % V == 0, "dtg_norm_bwd: channels");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int hw = dx->h * dx->w, n = dx->n, c = dx->c;
  int cg = 8 * V;
  while (c % cg != 0) cg >>= 1;
  const int splits = pick_splits(n, c, cg, hw);
  const dtg_plane p_dy2 = (dy2 && dy2->ptr) ? *dy2 : kNullPlane;
  const dtg_plane p_y = (y && y->ptr) ? *y : kNullPlane;
  const dtg_plane p_x = (x && x->ptr) ? *x : *dx;
  float* bnsum = partial;
  float* part = partial + 2 * c;
  float* kcoef = part + static_cast<size_t>(kMaxSplits) * n * c * 2;
  const bool need_reduce = mode != DTG_NORM_NONE || d_beta != nullptr || sums != nullptr;
  float* sums_buf = sums ? sums : kcoef + static_cast<size_t>(n) * c * 4;   // scratch when the caller wants none
  const int nc = n * c;
  // phase 3 / 4 split the per-channel parameter-gradient reduction (d_gamma / d_beta from the per-(n,c) sums) off the
  // data-gradient chain: 4 = everything but that reduction, 3 = that reduction alone (same arguments; the caller issues
  // it on another stream, after the phase-4 call).  Not for batch norm, whose channel sums feed dx.
  if (a->phase == 3) {
    DTG_REQUIRE(mode != DTG_NORM_COND_INSTANCE && mode != DTG_NORM_BATCH && (d_beta || d_gamma),
                "dtg_norm_bwd: phase 3 needs an instance / activation-only layer with parameter gradients");
    DTG_CHECK_CUDA(launch_k(norm_bwd_channel_kernel, (c + 31) / 32, 256, 0, stream, sums_buf, n, c, mode, bnsum, d_gamma, d_beta));
    return DTG_OK;
  }
  const bool defer_channel = a->phase == 4;
  DTG_REQUIRE(!defer_channel || mode != DTG_NORM_BATCH, "dtg_norm_bwd: phase 4 is not available for batch norm");
  dtg_norm_args a2 = *a;
  if (defer_channel) a2.phase = 0;
  a = &a2;
  if (mode == DTG_NORM_INSTANCE || mode == DTG_NORM_COND_INSTANCE || (mode == DTG_NORM_NONE && d_beta != nullptr)) {
    int rc = try_norm_bwd_lean(a, dy, dy2, y, &p_x, stats, gamma, sums_buf, part, dx, d_res, stream);
    if (rc == 1) rc = try_norm_bwd_tma(a, dy, dy2, y, &p_x, stats, gamma, sums_buf, dx, d_res, stream);
    if (rc == 1) rc = try_norm_bwd_fused(a, dy, dy2, y, &p_x, stats, gamma, sums_buf, dx, d_res, stream);
    if (rc < 0) return rc;
    DTG_REQUIRE(rc == 0 || dx->halo == 0, "dtg_norm_bwd: a dx plane with a halo needs the register-resident cluster kernel "
                                          "(instance / activation-only layer, <= 1024 pixels per slab, norm_impl 0 or 2)");
    if (rc == 0) {
      if (!defer_channel && mode != DTG_NORM_COND_INSTANCE && (d_beta || d_gamma)) {
        DTG_CHECK_CUDA(launch_k(norm_bwd_channel_kernel, (c + 31) / 32, 256, 0, stream, sums_buf, n, c, mode, bnsum, d_gamma, d_beta));
      }
      return DTG_OK;
    }
  }
  DTG_REQUIRE(dx->halo == 0, "dtg_norm_bwd: a dx plane with a halo is not supported for this mode / phase");
  if ((a->phase == 0 || a->phase == 1) && need_reduce) {
    dim3 grid(c / cg, n, splits);
    if (bf)
      DTG_CHECK_CUDA(launch_k(norm_bwd_reduce_kernel<__nv_bfloat16>, grid, kNormThreads, 0, stream, *dy, p_dy2, p_y, p_x, stats, mode, a->act, cg, splits, part));
    else
      DTG_CHECK_CUDA(launch_k(norm_bwd_reduce_kernel<float>, grid, kNormThreads, 0, stream, *dy, p_dy2, p_y, p_x, stats, mode, a->act, cg, splits, part));
    DTG_CHECK_CUDA(launch_k(norm_bwd_sums_kernel, (nc + 127) / 128, 128, 0, stream, part, splits, n, c, hw, mode, stats, gamma, sums_buf, kcoef));
    if (!defer_channel && mode != DTG_NORM_COND_INSTANCE && (mode == DTG_NORM_BATCH || d_beta || d_gamma)) {
      DTG_CHECK_CUDA(launch_k(norm_bwd_channel_kernel, (c + 31) / 32, 256, 0, stream, sums_buf, n, c, mode, bnsum, d_gamma, d_beta));
    }
  }
  if (a->phase == 1) return DTG_OK;
  if (mode == DTG_NORM_BATCH) {
    const int world = a->world_size > 0 ? a->world_size : 1;
    DTG_CHECK_CUDA(launch_k(norm_bwd_bn_kcoef_kernel, (nc + 127) / 128, 128, 0, stream, bnsum, n, c, hw, world, stats, gamma, kcoef));
  }
  const dtg_plane p_res = (d_res && d_res->ptr) ? *d_res : kNullPlane;
  const size_t total = static_cast<size_t>(n) * hw * (c / V);
  if (bf)
    DTG_CHECK_CUDA(launch_k(norm_bwd_apply_kernel<__nv_bfloat16>, elemwise_grid(total), kNormThreads, 0, stream, *dy, p_dy2, p_y, p_x, stats, kcoef, mode, a->act, *dx, p_res));
  else
    DTG_CHECK_CUDA(launch_k(norm_bwd_apply_kernel<float>, elemwise_grid(total), kNormThreads, 0, stream, *dy, p_dy2, p_y, p_x, stats, kcoef, mode, a->act, *dx, p_res));
  return DTG_OK;
}

extern "C" int dtg_cin_affine_fwd(const float* z, const float* ws, const float* bs, const float* wb, const float* bb,
                                  int n, int c, int nz, float* gamma, float* beta, void* stream) {
  DTG_REQUIRE(z && ws && bs && wb && bb && gamma && beta, "dtg_cin_affine_fwd: null argument");
  DTG_CHECK_CUDA(launch_k(cin_affine_fwd_kernel, (n * c + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream), z, ws, bs, wb, bb, n, c, nz, gamma, beta));
  return DTG_OK;
}

extern "C" int dtg_cin_affine_bwd(const float* z, const float* ws, const float* wb, const float* gamma, const float* beta,
                                  const float* sums, int n, int c, int nz, float* d_ws, float* d_bs, float* d_wb,
                                  float* d_bb, float* d_z, void* stream) {
  DTG_REQUIRE(z && ws && wb && gamma && beta && sums && d_ws && d_bs && d_wb && d_bb, "dtg_cin_affine_bwd: null argument");
  DTG_CHECK_CUDA(launch_k(cin_affine_bwd_kernel, c + (d_z ? n : 0), 128, 0, static_cast<cudaStream_t>(stream), z, ws, wb, gamma, beta, sums, n, c, nz, d_ws,
                                                                                           d_bs, d_wb, d_bb, d_z));
  return DTG_OK;
}
