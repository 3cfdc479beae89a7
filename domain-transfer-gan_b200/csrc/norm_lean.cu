// Two-phase (conditional) instance normalisation, forward and backward: plain streaming kernels.
//
// Phase 1 streams the inputs once and leaves per-(sample, pixel-chunk, channel) partial sums in a small workspace;
// phase 2 sums the partials of its sample in chunk order (deterministic, no atomics), derives the per-channel
// coefficients in shared memory and streams the inputs a second time -- 21 MB per tensor at batch 80, still resident
// in the 126 MB L2 when phase 2 follows phase 1 -- writing the outputs.  A thread owns one 16-byte channel vector of
// a pixel and ALL channels of a pixel are covered by adjacent threads, so every warp access is a run of whole pixels
// (512 contiguous bytes); there is no cluster, no barrier between the phases inside a kernel and no constraint on
// the grid, so 3-4 CTAs are resident per SM and the kernels behave like copies (the cluster-fused single-launch
// kernels in norm_fused.cu hold 16 warps per SM across a cluster barrier and reach a third of the HBM peak).
// Formulas: SURVEY.md 9.1 (modules.py:83-97,120-132 and their autograd); g = (fold(dy) + dy2) * act'(y).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "norm_common.cuh"

namespace dtg {

constexpr int kLT = 256;        // threads per CTA
constexpr int kLMaxC = 256;     // channels per pixel handled here (<= 32 vectors of 16 bytes: one warp spans whole pixels)
constexpr int kLMaxChunks = 32;

// (s1[V], s2[V]) of every thread -> per-channel sums of the CTA, written as float2 to out[channel]
template <int V>
__device__ __forceinline__ void lean_block_reduce(float (&s1)[V], float (&s2)[V], int nvc, int v, float (*red)[kLMaxC * 2],
                                                  float2* out) {
  const int tid = threadIdx.x;
  for (int off = nvc; off < 32; off <<= 1) {        // lanes of a warp that own the same channel vector are nvc apart
#pragma unroll
    for (int i = 0; i < V; ++i) {
      s1[i] += __shfl_xor_sync(0xffffffffu, s1[i], off);
      s2[i] += __shfl_xor_sync(0xffffffffu, s2[i], off);
    }
  }
  if ((tid & 31) < nvc) {
#pragma unroll
    for (int i = 0; i < V; ++i) *reinterpret_cast<float2*>(&red[tid >> 5][(v * V + i) * 2]) = make_float2(s1[i], s2[i]);
  }
  __syncthreads();
  if (tid < nvc * V) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < kLT / 32; ++w) {
      const float2 t = *reinterpret_cast<const float2*>(&red[w][tid * 2]);
      a += t.x;
      b += t.y;
    }
    out[tid] = make_float2(a, b);
  }
}

// ---- forward, phase 1: partial[n][chunk][c] = (sum(x-K), sum((x-K)^2)), K = x[n,0,0,c] ---------------------------------
template <typename T>
__global__ void __launch_bounds__(kLT, 4) lean_fwd_stats_kernel(dtg_plane x, int nvc, int chunk_px, float2* __restrict__ partial) {
  pdl_enter();
  constexpr int V = Vec<T>::N;
  __shared__ float red[kLT / 32][kLMaxC * 2];
  const int v = threadIdx.x % nvc, lane = threadIdx.x / nvc, lanes = kLT / nvc;
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int hw = x.h * x.w;
  const int p0 = chunk * chunk_px, p1 = min(hw, p0 + chunk_px);
  const size_t pitch = static_cast<size_t>(x.c) * sizeof(T);
  const uint8_t* xb = reinterpret_cast<const uint8_t*>(x.ptr) + static_cast<size_t>(n) * hw * pitch + v * 16;
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
  lean_block_reduce<V>(s1, s2, nvc, v, red, partial + (static_cast<size_t>(n) * gridDim.x + chunk) * x.c);
}

// ---- forward, phase 2: y = act(x*a + b (+ residual)), mirrored into the output halo ------------------------------------
template <typename T>
__global__ void __launch_bounds__(kLT, 4) lean_fwd_apply_kernel(dtg_plane x, dtg_plane res, dtg_plane out,
                                                                const float2* __restrict__ partial,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                float* __restrict__ stats, int mode, int act, float eps, int nvc,
                                                                int chunk_px) {
  pdl_enter();
  constexpr int V = Vec<T>::N;
  __shared__ float2 coef[kLMaxC];
  const int v = threadIdx.x % nvc, lane = threadIdx.x / nvc, lanes = kLT / nvc;
  const int n = blockIdx.y, chunk = blockIdx.x, chunks = gridDim.x;
  const int hw = x.h * x.w, C = x.c;
  const size_t pitch = static_cast<size_t>(C) * sizeof(T);
  if (threadIdx.x < C) {
    const int ch = threadIdx.x;
    float a1 = 0.f, a2 = 0.f;
    for (int k = 0; k < chunks; ++k) {
      const float2 t = partial[(static_cast<size_t>(n) * chunks + k) * C + ch];
      a1 += t.x;
      a2 += t.y;
    }
    const size_t first = static_cast<size_t>(n) * hw * C + ch;
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
    const size_t nc = static_cast<size_t>(n) * C + ch;
    const float ga = mode == DTG_NORM_COND_INSTANCE ? gamma[nc] : gamma[ch];
    const float be = mode == DTG_NORM_COND_INSTANCE ? beta[nc] : beta[ch];
    const float a = rstd * ga;
    coef[ch] = make_float2(a, be - mean * a);
    if (chunk == 0) {
      stats[nc * 2] = mean;
      stats[nc * 2 + 1] = rstd;
    }
  }
  __syncthreads();
  float ca[V], cb[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float2 ab = coef[v * V + i];
    ca[i] = ab.x;
    cb[i] = ab.y;
  }
  const float slope = act == DTG_ACT_RELU ? 0.f : (act == DTG_ACT_LRELU ? 0.2f : 1.f);
  const int p0 = chunk * chunk_px, p1 = min(hw, p0 + chunk_px);
  const uint8_t* xb = reinterpret_cast<const uint8_t*>(x.ptr) + static_cast<size_t>(n) * hw * pitch + v * 16;
  const int hr = res.ptr ? res.halo : 0, ho = out.halo;
  const uint8_t* rb = res.ptr ? reinterpret_cast<const uint8_t*>(res.ptr) +
                                    static_cast<size_t>(n) * (res.h + 2 * hr) * (res.w + 2 * hr) * pitch + v * 16
                              : nullptr;
  uint8_t* ob = reinterpret_cast<uint8_t*>(out.ptr) + static_cast<size_t>(n) * (out.h + 2 * ho) * (out.w + 2 * ho) * pitch + v * 16;
  const bool hal = hr > 0 || ho > 0;
  PixCur it;
  const int adv_q = lanes / x.w, adv_r = lanes - adv_q * x.w;
  it.init(p0 + lane, x.w, hr, ho);
#pragma unroll 2
  for (; it.p < p1;) {
    float f[V];
    Vec<T>::load(xb + it.p * pitch, f);
#pragma unroll
    for (int i = 0; i < V; ++i) f[i] = f[i] * ca[i] + cb[i];
    if (rb) {
      float t[V];
      Vec<T>::load(rb + static_cast<size_t>(it.o0) * pitch, t);
#pragma unroll
      for (int i = 0; i < V; ++i) f[i] += t[i];
    }
#pragma unroll
    for (int i = 0; i < V; ++i) f[i] = fmaxf(f[i], slope * f[i]);      // none (1) / ReLU (0) / LeakyReLU (0.2)
    Vec<T>::store(ob + static_cast<size_t>(it.o1) * pitch, f);
    if (ho > 0) {
      int hts[3], wts[3];
      const int nh = reflect_targets(it.py, out.h, ho, hts), nw = reflect_targets(it.px, out.w, ho, wts);
      if (nh * nw > 1) {
        const int wb = out.w + 2 * ho;
        for (int a = 0; a < nh; ++a)
          for (int q = 0; q < nw; ++q)
            if (a + q > 0) Vec<T>::store(ob + (static_cast<size_t>(hts[a] + ho) * wb + wts[q] + ho) * pitch, f);
      }
    }
    if (hal)
      it.template advance<true>(lanes, adv_q, adv_r, x.w, hr, ho);
    else
      it.template advance<false>(lanes, adv_q, adv_r, x.w, hr, ho);
  }
}

// ---- backward, phase 1: partial[n][chunk][c] = (sum g, sum g * xhat) -----------------------------------------------------
template <typename T, int ACT, bool HAL>
__global__ void __launch_bounds__(kLT, 3) lean_bwd_sums_kernel(dtg_plane dy, dtg_plane dy2, dtg_plane yp, dtg_plane x,
                                                               const float* __restrict__ stats, float2* __restrict__ partial,
                                                               int has_norm, int nvc, int chunk_px) {
  pdl_enter();
  constexpr int V = Vec<T>::N;
  __shared__ float red[kLT / 32][kLMaxC * 2];
  const int v = threadIdx.x % nvc, lane = threadIdx.x / nvc, lanes = kLT / nvc;
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int hw = dy.h * dy.w, C = dy.c;
  const int p0 = chunk * chunk_px, p1 = min(hw, p0 + chunk_px);
  const size_t pitch = static_cast<size_t>(C) * sizeof(T);
  const int c = v * V;
  const int hdy = dy.halo, hy = yp.ptr ? yp.halo : 0;
  const uint8_t* xb = has_norm ? reinterpret_cast<const uint8_t*>(x.ptr) + static_cast<size_t>(n) * hw * pitch + v * 16 : nullptr;
  const uint8_t* dyb = reinterpret_cast<const uint8_t*>(dy.ptr) + static_cast<size_t>(n) * (dy.h + 2 * hdy) * (dy.w + 2 * hdy) * pitch + v * 16;
  const uint8_t* yb = yp.ptr ? reinterpret_cast<const uint8_t*>(yp.ptr) + static_cast<size_t>(n) * (yp.h + 2 * hy) * (yp.w + 2 * hy) * pitch + v * 16
                             : nullptr;
  const uint8_t* dy2b = dy2.ptr ? reinterpret_cast<const uint8_t*>(dy2.ptr) + static_cast<size_t>(n) * hw * pitch + v * 16 : nullptr;
  float mean[V], rstd[V], s1[V], s2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    s1[i] = s2[i] = 0.f;
    mean[i] = rstd[i] = 0.f;
    if (has_norm) {
      const float2 mr = *reinterpret_cast<const float2*>(stats + (static_cast<size_t>(n) * C + c + i) * 2);
      mean[i] = mr.x;
      rstd[i] = mr.y;
    }
  }
  PixCur it;
  const int adv_q = lanes / dy.w, adv_r = lanes - adv_q * dy.w;
  it.init(p0 + lane, dy.w, hdy, hy);
#pragma unroll 2
  for (; it.p < p1; it.template advance<HAL>(lanes, adv_q, adv_r, dy.w, hdy, hy)) {
    float g[V];
    load_g_fast<T, ACT, HAL>(dy, dyb, dy2b, yb, it, n, c, g);
    if (has_norm) {
      float f[V];
      Vec<T>::load(xb + it.p * pitch, f);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        s1[i] += g[i];
        s2[i] += g[i] * ((f[i] - mean[i]) * rstd[i]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) s1[i] += g[i];
    }
  }
  lean_block_reduce<V>(s1, s2, nvc, v, red, partial + (static_cast<size_t>(n) * gridDim.x + chunk) * C);
}

// ---- backward, phase 2: dx = k0 * (g - A/m - xhat * B/d) = k0 * g + c1 + x * c2; d_res = g --------------------------------
template <typename T, int ACT, bool HAL>
__global__ void __launch_bounds__(kLT, 3) lean_bwd_apply_kernel(dtg_plane dy, dtg_plane dy2, dtg_plane yp, dtg_plane x,
                                                                const float* __restrict__ stats, const float* __restrict__ gamma,
                                                                const float2* __restrict__ partial, float* __restrict__ sums,
                                                                dtg_plane dx, dtg_plane dres, int mode, int nvc, int chunk_px) {
  pdl_enter();
  constexpr int V = Vec<T>::N;
  __shared__ float4 kco[kLMaxC];
  const int v = threadIdx.x % nvc, lane = threadIdx.x / nvc, lanes = kLT / nvc;
  const int n = blockIdx.y, chunk = blockIdx.x, chunks = gridDim.x;
  const int hw = dy.h * dy.w, C = dy.c;
  const bool has_norm = mode != DTG_NORM_NONE;
  if (threadIdx.x < C) {
    const int ch = threadIdx.x;
    float A = 0.f, B = 0.f;
    for (int k = 0; k < chunks; ++k) {
      const float2 t = partial[(static_cast<size_t>(n) * chunks + k) * C + ch];
      A += t.x;
      B += t.y;
    }
    const size_t nc = static_cast<size_t>(n) * C + ch;
    if (has_norm) {
      const float m = static_cast<float>(hw);
      const float d = mode == DTG_NORM_COND_INSTANCE ? m - 1.f : m;
      const float ga = mode == DTG_NORM_COND_INSTANCE ? gamma[nc] : gamma[ch];
      const float2 mr = *reinterpret_cast<const float2*>(stats + nc * 2);
      const float k0 = mr.y * ga;
      const float c2 = -k0 * mr.y * (B / d);
      kco[ch] = make_float4(k0, -k0 * (A / m) - mr.x * c2, c2, 0.f);
    } else {
      kco[ch] = make_float4(1.f, 0.f, 0.f, 0.f);
    }
    if (chunk == 0) {
      sums[nc * 2] = A;
      sums[nc * 2 + 1] = B;
    }
  }
  __syncthreads();
  const int p0 = chunk * chunk_px, p1 = min(hw, p0 + chunk_px);
  const size_t pitch = static_cast<size_t>(C) * sizeof(T);
  const int c = v * V;
  const int hdy = dy.halo, hy = yp.ptr ? yp.halo : 0;
  const uint8_t* xb = has_norm ? reinterpret_cast<const uint8_t*>(x.ptr) + static_cast<size_t>(n) * hw * pitch + v * 16 : nullptr;
  const uint8_t* dyb = reinterpret_cast<const uint8_t*>(dy.ptr) + static_cast<size_t>(n) * (dy.h + 2 * hdy) * (dy.w + 2 * hdy) * pitch + v * 16;
  const uint8_t* yb = yp.ptr ? reinterpret_cast<const uint8_t*>(yp.ptr) + static_cast<size_t>(n) * (yp.h + 2 * hy) * (yp.w + 2 * hy) * pitch + v * 16
                             : nullptr;
  const uint8_t* dy2b = dy2.ptr ? reinterpret_cast<const uint8_t*>(dy2.ptr) + static_cast<size_t>(n) * hw * pitch + v * 16 : nullptr;
  uint8_t* dxb = reinterpret_cast<uint8_t*>(dx.ptr) + static_cast<size_t>(n) * hw * pitch + v * 16;
  uint8_t* drb = dres.ptr ? reinterpret_cast<uint8_t*>(dres.ptr) + static_cast<size_t>(n) * hw * pitch + v * 16 : nullptr;
  float k0[V], c1[V], c2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float4 kk = kco[c + i];
    k0[i] = kk.x;
    c1[i] = kk.y;
    c2[i] = kk.z;
  }
  PixCur it;
  const int adv_q = lanes / dy.w, adv_r = lanes - adv_q * dy.w;
  it.init(p0 + lane, dy.w, hdy, hy);
#pragma unroll 2
  for (; it.p < p1; it.template advance<HAL>(lanes, adv_q, adv_r, dy.w, hdy, hy)) {
    float g[V];
    load_g_fast<T, ACT, HAL>(dy, dyb, dy2b, yb, it, n, c, g);
    if (drb) Vec<T>::store(drb + it.p * pitch, g);
    if (has_norm) {
      float f[V];
      Vec<T>::load(xb + it.p * pitch, f);
#pragma unroll
      for (int i = 0; i < V; ++i) g[i] = k0[i] * g[i] + (f[i] * c2[i] + c1[i]);
    }
    Vec<T>::store(dxb + it.p * pitch, g);
  }
}

struct LeanGeom {
  int nvc, chunks, chunk_px;
};

static bool lean_geom(const dtg_plane* x, LeanGeom* g) {
  const int es = elem_size(x->dtype);
  const int rowb = x->c * es;
  if (rowb % 16 != 0) return false;
  const int nvc = rowb / 16;
  if (nvc > 32 || (nvc & (nvc - 1)) != 0 || x->c > kLMaxC || x->n > 65535) return false;
  const int hw = x->h * x->w;
  if (hw < 2) return false;
  const int lanes = kLT / nvc;
  // enough CTAs for ~4 resident per SM, at least two sweeps of the CTA per chunk
  int chunks = (4 * 148 + x->n - 1) / x->n;
  chunks = std::min(chunks, std::max(1, hw / (2 * lanes)));
  chunks = std::max(1, std::min(chunks, kLMaxChunks));
  int chunk_px = (hw + chunks - 1) / chunks;
  chunk_px = (chunk_px + lanes - 1) / lanes * lanes;
  g->nvc = nvc;
  g->chunk_px = chunk_px;
  g->chunks = (hw + chunk_px - 1) / chunk_px;
  return true;
}

static const dtg_plane kNullL = {nullptr, 0, 0, 0, 0, 0, 0};

// returns DTG_OK after launching, 1 when the geometry is not handled (caller falls back), < 0 on errors.
// `partial`: >= n * chunks * c float2 (dtg_norm_workspace_bytes covers 32 chunks)
int try_norm_fwd_lean(const dtg_norm_args* a, const dtg_plane* x, const dtg_plane* residual, const float* gamma,
                      const float* beta, float* stats, float* partial, const dtg_plane* out, cudaStream_t stream) {
  if ((norm_impl() != 2 && norm_impl() != 3) || a->phase != 0 || (a->mode != DTG_NORM_INSTANCE && a->mode != DTG_NORM_COND_INSTANCE)) return 1;
  if (a->act == DTG_ACT_TANH || x->halo != 0) return 1;
  LeanGeom g;
  if (!lean_geom(x, &g)) return 1;
  const dtg_plane res = (residual && residual->ptr) ? *residual : kNullL;
  float2* part = reinterpret_cast<float2*>(partial);
  const dim3 grid(g.chunks, x->n, 1);
  if (x->dtype == DTG_BF16) {
    DTG_CHECK_CUDA(launch_k(lean_fwd_stats_kernel<__nv_bfloat16>, grid, kLT, 0, stream, *x, g.nvc, g.chunk_px, part));
    DTG_CHECK_CUDA(launch_k(lean_fwd_apply_kernel<__nv_bfloat16>, grid, kLT, 0, stream, *x, res, *out, part, gamma, beta, stats,
                            static_cast<int>(a->mode), static_cast<int>(a->act), a->eps, g.nvc, g.chunk_px));
  } else {
    DTG_CHECK_CUDA(launch_k(lean_fwd_stats_kernel<float>, grid, kLT, 0, stream, *x, g.nvc, g.chunk_px, part));
    DTG_CHECK_CUDA(launch_k(lean_fwd_apply_kernel<float>, grid, kLT, 0, stream, *x, res, *out, part, gamma, beta, stats,
                            static_cast<int>(a->mode), static_cast<int>(a->act), a->eps, g.nvc, g.chunk_px));
  }
  return DTG_OK;
}

int try_norm_bwd_lean(const dtg_norm_args* a, const dtg_plane* dy, const dtg_plane* dy2, const dtg_plane* y,
                      const dtg_plane* x, const float* stats, const float* gamma, float* sums, float* partial,
                      const dtg_plane* dx, const dtg_plane* d_res, cudaStream_t stream) {
  if (norm_impl() != 3 || a->phase != 0 || a->mode == DTG_NORM_BATCH) return 1;
  const bool has_norm = a->mode != DTG_NORM_NONE;
  const dtg_plane p_dy2 = (dy2 && dy2->ptr) ? *dy2 : kNullL;
  const dtg_plane p_y = (a->act != DTG_ACT_NONE && y && y->ptr) ? *y : kNullL;
  const dtg_plane p_res = (d_res && d_res->ptr) ? *d_res : kNullL;
  if (a->act != DTG_ACT_NONE && !p_y.ptr) return 1;
  if (dx->halo != 0 || (p_res.ptr && p_res.halo != 0) || (p_dy2.ptr && p_dy2.halo != 0)) return 1;
  if (dy->halo > 1 || (dy->halo == 1 && (dy->h < 4 || dy->w < 4))) return 1;      // load_g_fast's fast reflect-fold
  LeanGeom g;
  if (!lean_geom(dx, &g)) return 1;
  DTG_REQUIRE(dy->c == dx->c && (!has_norm || x->c == dx->c) && (!p_y.ptr || p_y.c == dx->c) && (!p_dy2.ptr || p_dy2.c == dx->c) &&
                  (!p_res.ptr || p_res.c == dx->c),
              "norm_bwd_lean: channel mismatch");
  const bool hal = dy->halo > 0 || (p_y.ptr && p_y.halo > 0);
  float2* part = reinterpret_cast<float2*>(partial);
  const dim3 grid(g.chunks, dx->n, 1);
  const dtg_plane px = has_norm ? *x : *dx;
#define DTG_LEAN_BWD(TT, AA, HH)                                                                                              \
  do {                                                                                                                        \
    DTG_CHECK_CUDA(launch_k(lean_bwd_sums_kernel<TT, AA, HH>, grid, kLT, 0, stream, *dy, p_dy2, p_y, px, stats, part,         \
                            has_norm ? 1 : 0, g.nvc, g.chunk_px));                                                            \
    DTG_CHECK_CUDA(launch_k(lean_bwd_apply_kernel<TT, AA, HH>, grid, kLT, 0, stream, *dy, p_dy2, p_y, px, stats, gamma, part, \
                            sums, *dx, p_res, static_cast<int>(a->mode), g.nvc, g.chunk_px));                                 \
    return DTG_OK;                                                                                                            \
  } while (0)
#define DTG_LEAN_BWD2(TT, AA) \
  do {                        \
    if (hal)                  \
      DTG_LEAN_BWD(TT, AA, true); \
    else                      \
      DTG_LEAN_BWD(TT, AA, false); \
  } while (0)
  if (dx->dtype == DTG_BF16) {
    if (a->act == DTG_ACT_RELU) DTG_LEAN_BWD2(__nv_bfloat16, DTG_ACT_RELU);
    if (a->act == DTG_ACT_LRELU) DTG_LEAN_BWD2(__nv_bfloat16, DTG_ACT_LRELU);
    DTG_LEAN_BWD2(__nv_bfloat16, DTG_ACT_NONE);
  }
  if (a->act == DTG_ACT_RELU) DTG_LEAN_BWD2(float, DTG_ACT_RELU);
  if (a->act == DTG_ACT_LRELU) DTG_LEAN_BWD2(float, DTG_ACT_LRELU);
  DTG_LEAN_BWD2(float, DTG_ACT_NONE);
#undef DTG_LEAN_BWD2
#undef DTG_LEAN_BWD
}

}  // namespace dtg
