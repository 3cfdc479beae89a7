// TMA-staged (conditional) instance normalisation, forward and backward, on thread-block clusters.
//
// A "slab" is one sample x one channel block (128 / 64 / 32 bytes of channels) x all pixels; a cluster of CS CTAs
// owns a slab at a time and splits its rows.  Every input of the CTA's row range (forward: x, residual; backward:
// dy, dy2, y, x) is brought into shared memory by ONE TMA box each (cp.async.bulk.tensor, mbarrier complete_tx) and
// the results leave through TMA stores, so the memory pipeline holds a whole work item per CTA without costing a
// register, 2-3 CTAs are resident per SM (one loads while another reduces / applies) and HBM traffic is exactly one
// read of every input plus one write of every output.  Pass 1 reduces the per-channel sums from shared memory (warp
// shuffles -> per-warp partials -> fixed-order block sum -> pushed into every peer's shared memory through DSMEM ->
// summed in rank order: deterministic, no atomics); pass 2 re-reads the still-resident data and writes the output in
// place.  All shared-memory traffic uses 32-bit addresses and 128-bit ld/st; the per-thread pixel geometry is
// computed once per kernel (no divisions in the item loop).
// Formulas: SURVEY.md 9.1 (modules.py:83-97,120-132 and their autograd); g = (fold(dy) + dy2) * act'(y).
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>

#include "common.cuh"
#include "conv_epilogue.cuh"
#include "norm_common.cuh"

namespace cg = cooperative_groups;

namespace dtg {

constexpr int kTN = 256;            // threads per CTA
constexpr int kTSlab = 64;          // max channels per slab
constexpr int kTMaxCluster = 8;
constexpr int kTSetMax = 96 * 1024;     // bytes of one work item's buffers: 2 CTAs / SM
constexpr int kTSetGood = 58 * 1024;    // 3 CTAs / SM
constexpr uint32_t kTInvalid = 0xFFFFFFFFu;

__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 r;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a) : "memory");
  return r;
}
__device__ __forceinline__ void sts128(uint32_t a, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
template <typename T>
__device__ __forceinline__ uint4 pack_vec(const float (&f)[Vec<T>::N]) {
  if constexpr (sizeof(T) == 2) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  } else {
    return make_uint4(__float_as_uint(round_tf32(f[0])), __float_as_uint(round_tf32(f[1])), __float_as_uint(round_tf32(f[2])),
                      __float_as_uint(round_tf32(f[3])));
  }
}
template <typename T>
__device__ __forceinline__ uint4 raw_vec(const float (&f)[Vec<T>::N]) {     // fp32 planes: unrounded
  if constexpr (sizeof(T) == 2) {
    return pack_vec<T>(f);
  } else {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
  }
}

// Fixed-order reduction of per-thread (s1[V], s2[V]) over the pixel lanes of the CTA and the CTAs of the cluster:
// xor tree over the lanes of a warp that own the same channel vector (nv apart) -> per-warp partials -> block sum in
// warp order -> pushed into slot `rank` of every peer's allpart[] through DSMEM -> after ONE cluster barrier every
// CTA sums the cs partials locally in rank order.  Returns the totals for channel `tid` (valid for tid < nv * V).
template <int V>
__device__ __forceinline__ float2 cluster_reduce2(float (&s1)[V], float (&s2)[V], int nv, int v, int tid,
                                                  float (*wred)[kTSlab * 2], float (*allpart)[kTSlab * 2],
                                                  cg::cluster_group& cl, int cs, int rank) {
  for (int off = nv; off < 32; off <<= 1) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      s1[i] += __shfl_xor_sync(0xffffffffu, s1[i], off);
      s2[i] += __shfl_xor_sync(0xffffffffu, s2[i], off);
    }
  }
  if ((tid & 31) < nv) {
#pragma unroll
    for (int i = 0; i < V; ++i) *reinterpret_cast<float2*>(&wred[tid >> 5][(v * V + i) * 2]) = make_float2(s1[i], s2[i]);
  }
  __syncthreads();
  const int slab_ch = nv * V;
  if (tid < slab_ch) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < kTN / 32; ++w) {
      const float2 x = *reinterpret_cast<const float2*>(&wred[w][tid * 2]);
      a += x.x;
      b += x.y;
    }
    float2* mine = reinterpret_cast<float2*>(&allpart[rank][tid * 2]);
    if (cs > 1) {
      for (int r = 0; r < cs; ++r) *cl.map_shared_rank(mine, r) = make_float2(a, b);
    } else {
      *mine = make_float2(a, b);
    }
  }
  if (cs > 1)
    cl.sync();
  else
    __syncthreads();
  float A = 0.f, B = 0.f;
  if (tid < slab_ch) {
    for (int r = 0; r < cs; ++r) {
      const float2 x = *reinterpret_cast<const float2*>(&allpart[r][tid * 2]);
      A += x.x;
      B += x.y;
    }
  }
  return make_float2(A, B);
}

struct NormTParams {
  CUtensorMap tm_dy, tm_dy2, tm_y, tm_x, tm_dx, tm_dres;
  const float* stats;
  const float* gamma;
  float* sums;
  int has_dy2, has_y, has_x, has_dres;
  int mode;
  int H, W, C, N;
  int nv;                   // 16-byte vectors per pixel of a slab
  int cblocks, items, nclusters;
  int cs, R;                // cluster size, rows per CTA
  int dy_pad;               // dy is a halo-1 plane, loaded as a box with its halo columns and one / two extra rows
  int y_h;                  // halo of the y plane (only its interior is loaded)
  int off_b1, off_y, off_x, set_bytes;
  unsigned tx_bytes;
};

// PPT: pixels per thread (upper bound); PAD: dy is a padded (halo 1) box whose ring is folded back
template <typename T, int ACT, int PPT, bool PAD>
__global__ void __launch_bounds__(kTN, 3) norm_bwd_tma_kernel(const __grid_constant__ NormTParams p) {
  constexpr int V = Vec<T>::N;
  extern __shared__ uint8_t smem_raw[];
  __shared__ float wred[kTN / 32][kTSlab * 2];
  __shared__ float allpart[2][kTMaxCluster][kTSlab * 2];
  __shared__ float4 kco[kTSlab];
  __shared__ __align__(8) uint64_t bar_full;
  pdl_trigger();
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sb = smem_u32(smem);
  cg::cluster_group cl = cg::this_cluster();
  const int cs = p.cs;
  const int rank = cs > 1 ? static_cast<int>(cl.block_rank()) : 0;
  const int cluster_id = blockIdx.x / cs;
  const int tid = threadIdx.x;
  const int nv = p.nv, slab_ch = nv * V;
  const int v = tid % nv, lane = tid / nv, lanes = kTN / nv;
  const int r0 = rank * p.R;
  const int npix = p.R * p.W;
  const bool has_norm = p.has_x != 0;

  // per-thread pixel geometry, identical for every work item: byte offsets inside the dense buffers (od), inside the
  // dy box (oy) and the reflect-fold neighbours (byte deltas; 0 = none)
  uint32_t od[PPT], oy[PAD ? PPT : 1];
  int fr[PAD ? PPT : 1], fc[PAD ? PPT : 1];
  {
    const int dy_w = p.W + 2;
    const int row0 = rank == 0 ? 1 : 0;           // local box row of the CTA's first interior row
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      const int pl = lane + j * lanes;
      const int ly = pl / p.W, lx = pl - ly * p.W;
      const bool ok = pl < npix && r0 + ly < p.H;
      od[j] = ok ? static_cast<uint32_t>((pl * nv + v) * 16) : kTInvalid;
      if (PAD) {
        const int y = r0 + ly;
        const int cell = (ly + row0) * dy_w + lx + 1;
        oy[j] = static_cast<uint32_t>((cell * nv + v) * 16);
        // reflection-pad(1) backward: row 1 also receives the halo row above row 0, row H-2 the one below row H-1
        fr[j] = (y == 1 ? -2 * dy_w : (y == p.H - 2 ? 2 * dy_w : 0)) * nv * 16;
        fc[j] = (lx == 1 ? -2 : (lx == p.W - 2 ? 2 : 0)) * nv * 16;
      }
    }
  }
  const uint32_t goff = PAD ? static_cast<uint32_t>(p.off_b1) : 0u;     // where g lives: dy's buffer in place when dense

  if (tid == 0) {
    tma_prefetch_desc(&p.tm_dy);
    tma_prefetch_desc(&p.tm_dx);
    if (p.has_dy2) tma_prefetch_desc(&p.tm_dy2);
    if (p.has_y) tma_prefetch_desc(&p.tm_y);
    if (p.has_x) tma_prefetch_desc(&p.tm_x);
    if (p.has_dres) tma_prefetch_desc(&p.tm_dres);
    mbar_init(&bar_full, 1);
    mbar_fence_init();
  }
  __syncthreads();
  pdl_wait();

  int k = 0;
  for (int item = cluster_id; item < p.items; item += p.nclusters, ++k) {
    const int n = item / p.cblocks, c0 = (item - n * p.cblocks) * slab_ch;
    if (tid == 0) {     // the buffers are free: the previous item's stores have been read out (end of the loop body)
      mbar_expect_tx(&bar_full, p.tx_bytes);
      if (PAD)
        tma_load_4d(smem, &p.tm_dy, &bar_full, c0, 0, rank == 0 ? 0 : r0 + 1, n);
      else
        tma_load_4d(smem, &p.tm_dy, &bar_full, c0, 0, r0, n);
      if (p.has_dy2) tma_load_4d(smem + p.off_b1, &p.tm_dy2, &bar_full, c0, 0, r0, n);
      if (p.has_y) tma_load_4d(smem + p.off_y, &p.tm_y, &bar_full, c0, p.y_h, r0 + p.y_h, n);
      if (p.has_x) tma_load_4d(smem + p.off_x, &p.tm_x, &bar_full, c0, 0, r0, n);
    }
    float mean[V], rstd[V], s1[V], s2[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      s1[i] = s2[i] = 0.f;
      mean[i] = rstd[i] = 0.f;
      if (has_norm) {
        const float2 mr = *reinterpret_cast<const float2*>(p.stats + (static_cast<size_t>(n) * p.C + c0 + v * V + i) * 2);
        mean[i] = mr.x;
        rstd[i] = mr.y;
      }
    }
    mbar_wait(&bar_full, static_cast<uint32_t>(k) & 1u);

    // ---- pass 1: g -> its buffer; A = sum g, B = sum g * xhat over the CTA's rows
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      if (od[j] == kTInvalid) continue;
      float g[V];
      if (PAD) {
        const uint32_t a = sb + oy[j];
        Vec<T>::unpack(lds128(a), g);
        if (fr[j] != 0) {
          float t[V];
          Vec<T>::unpack(lds128(a + fr[j]), t);
#pragma unroll
          for (int i = 0; i < V; ++i) g[i] += t[i];
        }
        if (fc[j] != 0) {
          float t[V];
          Vec<T>::unpack(lds128(a + fc[j]), t);
#pragma unroll
          for (int i = 0; i < V; ++i) g[i] += t[i];
          if (fr[j] != 0) {
            Vec<T>::unpack(lds128(a + fr[j] + fc[j]), t);
#pragma unroll
            for (int i = 0; i < V; ++i) g[i] += t[i];
          }
        }
      } else {
        Vec<T>::unpack(lds128(sb + od[j]), g);
      }
      if (p.has_dy2) {
        float t[V];
        Vec<T>::unpack(lds128(sb + p.off_b1 + od[j]), t);
#pragma unroll
        for (int i = 0; i < V; ++i) g[i] += t[i];
      }
      if (ACT != DTG_ACT_NONE) {
        float t[V];
        Vec<T>::unpack(lds128(sb + p.off_y + od[j]), t);
#pragma unroll
        for (int i = 0; i < V; ++i) g[i] = t[i] > 0.f ? g[i] : (ACT == DTG_ACT_LRELU ? 0.2f * g[i] : 0.f);
      }
      // bf16 planes keep g in its stored precision (it IS the residual-branch gradient); fp32 planes keep it unrounded
      // until pass 2 (activation-only layers store the final value right away)
      sts128(sb + goff + od[j], has_norm ? raw_vec<T>(g) : pack_vec<T>(g));
      if (has_norm) {
        float f[V];
        Vec<T>::unpack(lds128(sb + p.off_x + od[j]), f);
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
    const float2 tot = cluster_reduce2<V>(s1, s2, nv, v, tid, wred, allpart[k & 1], cl, cs, rank);
    if (tid < slab_ch) {
      const float A = tot.x, B = tot.y;
      const int ch = c0 + tid;
      const size_t nc = static_cast<size_t>(n) * p.C + ch;
      if (has_norm) {
        const float m = static_cast<float>(p.H * p.W);
        const float d = p.mode == DTG_NORM_COND_INSTANCE ? m - 1.f : m;
        const float ga = p.mode == DTG_NORM_COND_INSTANCE ? p.gamma[nc] : p.gamma[ch];
        const float2 mr = *reinterpret_cast<const float2*>(p.stats + nc * 2);
        // dx = k0 * (g - A/m - xhat * B/d) = k0 * g + c1 + x * c2
        const float k0 = mr.y * ga;
        const float c2 = -k0 * mr.y * (B / d);
        kco[tid] = make_float4(k0, -k0 * (A / m) - mr.x * c2, c2, 0.f);
      }
      if (rank == 0) {
        p.sums[nc * 2] = A;
        p.sums[nc * 2 + 1] = B;
      }
    }
    uint32_t outb = sb + goff;        // activation-only layers: dx = g, already in place
    if (has_norm) {
      __syncthreads();
      // ---- pass 2: dx in place of x; fp32 planes round the residual-branch gradient now
      float k0[V], c1[V], c2[V];
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float4 kk = kco[v * V + i];
        k0[i] = kk.x;
        c1[i] = kk.y;
        c2[i] = kk.z;
      }
      outb = sb + p.off_x;
#pragma unroll
      for (int j = 0; j < PPT; ++j) {
        if (od[j] == kTInvalid) continue;
        float g[V], f[V];
        Vec<T>::unpack(lds128(sb + goff + od[j]), g);
        Vec<T>::unpack(lds128(outb + od[j]), f);
        if (sizeof(T) == 4 && p.has_dres) sts128(sb + goff + od[j], pack_vec<T>(g));
#pragma unroll
        for (int i = 0; i < V; ++i) g[i] = k0[i] * g[i] + (f[i] * c2[i] + c1[i]);
        sts128(outb + od[j], pack_vec<T>(g));
      }
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tma_store_4d(&p.tm_dx, smem + (outb - sb), c0, 0, r0, n);
      if (p.has_dres) tma_store_4d(&p.tm_dres, smem + goff, c0, 0, r0, n);
      bulk_commit();
      bulk_wait_read<0>();          // the stores have finished reading the buffers: they may be refilled
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// forward: stats (shifted sums, K = first pixel of the sample) -> y = act(x*a + b (+ residual)), same structure.
// The interior of the output leaves through a TMA store; mirrored copies into the output's reflect halo are a few
// direct 16-byte stores from the border pixels' registers.
// ---------------------------------------------------------------------------------------------------------------
struct NormFParams {
  CUtensorMap tm_x, tm_res, tm_out;
  const void* xptr;
  const float* gamma;
  const float* beta;
  float* stats;
  dtg_plane out;
  int has_res, res_h;
  int mode, act;
  float eps;
  int H, W, C, N;
  int nv, cblocks, items, nclusters, cs, R;
  int off_res, set_bytes;
  unsigned tx_bytes;
};

template <typename T, int PPT>
__global__ void __launch_bounds__(kTN, 3) norm_fwd_tma_kernel(const __grid_constant__ NormFParams p) {
  constexpr int V = Vec<T>::N;
  extern __shared__ uint8_t smem_raw[];
  __shared__ float wred[kTN / 32][kTSlab * 2];
  __shared__ float allpart[2][kTMaxCluster][kTSlab * 2];
  __shared__ float2 coef[kTSlab];
  __shared__ __align__(8) uint64_t bar_full;
  pdl_trigger();
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sb = smem_u32(smem);
  cg::cluster_group cl = cg::this_cluster();
  const int cs = p.cs;
  const int rank = cs > 1 ? static_cast<int>(cl.block_rank()) : 0;
  const int cluster_id = blockIdx.x / cs;
  const int tid = threadIdx.x;
  const int nv = p.nv, slab_ch = nv * V;
  const int v = tid % nv, lane = tid / nv, lanes = kTN / nv;
  const int r0 = rank * p.R;
  const int npix = p.R * p.W;

  uint32_t od[PPT], pyx[PPT];      // dense byte offset; (image row << 16) | column
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const int pl = lane + j * lanes;
    const int ly = pl / p.W, lx = pl - ly * p.W;
    const bool ok = pl < npix && r0 + ly < p.H;
    od[j] = ok ? static_cast<uint32_t>((pl * nv + v) * 16) : kTInvalid;
    pyx[j] = (static_cast<uint32_t>(r0 + ly) << 16) | static_cast<uint32_t>(lx);
  }

  if (tid == 0) {
    tma_prefetch_desc(&p.tm_x);
    tma_prefetch_desc(&p.tm_out);
    if (p.has_res) tma_prefetch_desc(&p.tm_res);
    mbar_init(&bar_full, 1);
    mbar_fence_init();
  }
  __syncthreads();
  pdl_wait();

  int k = 0;
  for (int item = cluster_id; item < p.items; item += p.nclusters, ++k) {
    const int n = item / p.cblocks, c0 = (item - n * p.cblocks) * slab_ch;
    if (tid == 0) {
      mbar_expect_tx(&bar_full, p.tx_bytes);
      tma_load_4d(smem, &p.tm_x, &bar_full, c0, 0, r0, n);
      if (p.has_res) tma_load_4d(smem + p.off_res, &p.tm_res, &bar_full, c0, p.res_h, r0 + p.res_h, n);
    }
    // shift K = first pixel of the sample (identical in every CTA of the cluster; an L2 hit)
    float K[V], s1[V], s2[V];
    Vec<T>::load(reinterpret_cast<const uint8_t*>(p.xptr) +
                     (static_cast<size_t>(n) * p.H * p.W * p.C + c0 + v * V) * sizeof(T), K);
#pragma unroll
    for (int i = 0; i < V; ++i) s1[i] = s2[i] = 0.f;
    mbar_wait(&bar_full, static_cast<uint32_t>(k) & 1u);
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      if (od[j] == kTInvalid) continue;
      float f[V];
      Vec<T>::unpack(lds128(sb + od[j]), f);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float d = f[i] - K[i];
        s1[i] += d;
        s2[i] += d * d;
      }
    }
    const float2 tot = cluster_reduce2<V>(s1, s2, nv, v, tid, wred, allpart[k & 1], cl, cs, rank);
    if (tid < slab_ch) {
      const float a1 = tot.x, a2 = tot.y;
      const int ch = c0 + tid;
      const size_t first = static_cast<size_t>(n) * p.H * p.W * p.C + ch;
      float Kc;
      if constexpr (sizeof(T) == 2)
        Kc = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.xptr)[first]);
      else
        Kc = reinterpret_cast<const float*>(p.xptr)[first];
      const float m = static_cast<float>(p.H * p.W);
      const float mean = Kc + a1 / m;
      const float d = p.mode == DTG_NORM_COND_INSTANCE ? m - 1.f : m;
      float var = (a2 - a1 * a1 / m) / d;
      var = var < 0.f ? 0.f : var;
      const float rstd = rsqrtf(var + p.eps);
      const size_t nc = static_cast<size_t>(n) * p.C + ch;
      const float ga = p.mode == DTG_NORM_COND_INSTANCE ? p.gamma[nc] : p.gamma[ch];
      const float be = p.mode == DTG_NORM_COND_INSTANCE ? p.beta[nc] : p.beta[ch];
      const float a = rstd * ga;
      coef[tid] = make_float2(a, be - mean * a);
      if (rank == 0) {
        p.stats[nc * 2] = mean;
        p.stats[nc * 2 + 1] = rstd;
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
    const float slope = p.act == DTG_ACT_RELU ? 0.f : (p.act == DTG_ACT_LRELU ? 0.2f : 1.f);
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      if (od[j] == kTInvalid) continue;
      float f[V];
      Vec<T>::unpack(lds128(sb + od[j]), f);
#pragma unroll
      for (int i = 0; i < V; ++i) f[i] = f[i] * ca[i] + cb[i];
      if (p.has_res) {
        float t[V];
        Vec<T>::unpack(lds128(sb + p.off_res + od[j]), t);
#pragma unroll
        for (int i = 0; i < V; ++i) f[i] += t[i];
      }
#pragma unroll
      for (int i = 0; i < V; ++i) f[i] = fmaxf(f[i], slope * f[i]);     // none (1) / ReLU (0) / LeakyReLU (0.2)
      const uint4 o = pack_vec<T>(f);
      sts128(sb + od[j], o);
      if (p.out.halo > 0) {
        const int py = static_cast<int>(pyx[j] >> 16), px = static_cast<int>(pyx[j] & 0xFFFFu);
        int hts[3], wts[3];
        const int nh = reflect_targets(py, p.out.h, p.out.halo, hts), nw = reflect_targets(px, p.out.w, p.out.halo, wts);
        if (nh * nw > 1) {
          uint8_t* ob = reinterpret_cast<uint8_t*>(p.out.ptr);
          for (int a = 0; a < nh; ++a)
            for (int q = 0; q < nw; ++q)
              if (a + q > 0)
                *reinterpret_cast<uint4*>(ob + (plane_pix(p.out, n, hts[a], wts[q]) * p.out.c + c0 + v * V) * sizeof(T)) = o;
        }
      }
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tma_store_4d(&p.tm_out, smem, c0, p.out.halo, r0 + p.out.halo, n);
      bulk_commit();
      bulk_wait_read<0>();
    }
  }
}

struct TGeom {
  int slabb, cs, R, dense, dyb, set_bytes, need_b1, ppt;
};

// nbuf_dense: dense buffers besides the first one; pad: the first buffer is a halo-1 box
static bool tma_geom(const dtg_plane* dy, int nbuf_dense, bool pad, TGeom* g) {
  const int es = elem_size(dy->dtype);
  const int rowb = dy->c * es;
  const int H = dy->h, W = dy->w, N = dy->n;
  if (pad && (H < 4 || W < 4)) return false;
  if (W > 254 || H < 1) return false;
  // widest slab (DRAM-friendly rows); per slab width the smallest cluster whose work item fits the 3-CTA budget, else
  // the 2-CTA budget, with at most 8 pixels per thread
  for (int slabb = 128; slabb >= 32; slabb >>= 1) {
    if (rowb % slabb != 0) continue;
    if (slabb == 32 && rowb != 32) continue;
    const int lanes = kTN / (slabb / 16);
    const int items = N * (rowb / slabb);
    for (int pass = 0; pass < 2; ++pass) {
      const int budget = pass == 0 ? kTSetGood : kTSetMax;
      for (int cs = 1; cs <= kTMaxCluster; cs *= 2) {
        const int R = (H + cs - 1) / cs;
        if ((cs - 1) * R >= H) break;                         // a CTA without rows
        if (pad && (H % cs != 0 || R < 2)) continue;
        if (R + 2 > 256 || R * W > lanes * 8) continue;
        if (pass == 0 && items * cs < 2 * 148 && cs < kTMaxCluster) continue;     // prefer enough CTAs to fill the GPU
        const int dense = (R * W * slabb + 127) & ~127;
        const int dyb = pad ? (((R + (cs == 1 ? 2 : 1)) * (W + 2) * slabb + 127) & ~127) : dense;
        const int set = dyb + nbuf_dense * dense;
        if (set > budget) continue;
        g->slabb = slabb;
        g->cs = cs;
        g->R = R;
        g->dense = dense;
        g->dyb = dyb;
        g->set_bytes = (set + 1023) & ~1023;
        g->ppt = R * W <= lanes * 4 ? 4 : 8;
        return true;
      }
    }
  }
  return false;
}

static int plane_map(CUtensorMap* m, const dtg_plane* pl, int box_c, int box_w, int box_h) {
  const int es = elem_size(pl->dtype);
  const int Hb = pl->h + 2 * pl->halo, Wb = pl->w + 2 * pl->halo;
  uint64_t dims[4] = {static_cast<uint64_t>(pl->c), static_cast<uint64_t>(Wb), static_cast<uint64_t>(Hb), static_cast<uint64_t>(pl->n)};
  uint64_t strides[3] = {static_cast<uint64_t>(pl->c) * es, static_cast<uint64_t>(Wb) * pl->c * es,
                         static_cast<uint64_t>(Hb) * Wb * pl->c * es};
  uint32_t box[4] = {static_cast<uint32_t>(box_c), static_cast<uint32_t>(box_w), static_cast<uint32_t>(box_h), 1u};
  return encode_tiled(m, pl->dtype, 4, pl->ptr, dims, strides, box, 0);
}

template <typename K, typename P>
static int launch_tma_norm(K kernel, const P& p0, size_t smem, cudaStream_t stream) {
  P p = p0;
  static std::mutex mu;
  static std::map<std::pair<const void*, long long>, int> cache;     // (kernel, cs | smem) -> co-resident clusters
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.blockDim = dim3(kTN, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = p.cs;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  int ncl = 0;
  {
    std::lock_guard<std::mutex> lk(mu);
    const auto key = std::make_pair(reinterpret_cast<const void*>(kernel), (static_cast<long long>(p.cs) << 32) | static_cast<long long>(smem));
    auto it = cache.find(key);
    if (it == cache.end()) {
      DTG_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTSetMax + 2048));
      cfg.gridDim = dim3(p.cs * 148, 1, 1);
      cfg.numAttrs = 1;
      DTG_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&ncl, kernel, &cfg));
      if (ncl < 1) ncl = 1;
      static const bool dbg_occ = getenv("DTG_DEBUG_OCC") != nullptr;
      if (dbg_occ) fprintf(stderr, "[dtg] tma norm: cluster %d, smem %zu: max active clusters %d (= %d CTAs)\n", p.cs, smem, ncl, ncl * p.cs);
      cache[key] = ncl;
    } else {
      ncl = it->second;
    }
  }
  p.nclusters = std::min(p.items, ncl);
  cfg.gridDim = dim3(p.nclusters * p.cs, 1, 1);
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  DTG_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, p));
  DTG_LAUNCHED();
  return DTG_OK;
}

// returns DTG_OK after launching, 1 when the geometry is not handled (caller falls back), < 0 on errors
int try_norm_bwd_tma(const dtg_norm_args* a, const dtg_plane* dy, const dtg_plane* dy2, const dtg_plane* y,
                     const dtg_plane* x, const float* stats, const float* gamma, float* sums, const dtg_plane* dx,
                     const dtg_plane* d_res, cudaStream_t stream) {
  if (norm_impl() != 1 || a->phase != 0 || a->mode == DTG_NORM_BATCH) return 1;
  const bool has_norm = a->mode != DTG_NORM_NONE;
  const bool has_dy2 = dy2 && dy2->ptr, has_y = a->act != DTG_ACT_NONE, has_dres = d_res && d_res->ptr;
  if (!has_norm && !has_y) return 1;
  if (has_y && !(y && y->ptr)) return 1;
  if (dy->n > 65535 || dy->halo > 1 || dx->halo != 0 || (has_dres && d_res->halo != 0) || (has_dy2 && dy2->halo != 0)) return 1;
  if (has_dres && !has_norm) return 1;
  const bool pad = dy->halo == 1;
  // buffers: dy (g in place when dense) | b1 = dy2 / g / d_res (present when dy is padded or dy2 exists) | y | x
  const int need_b1 = (pad || has_dy2) ? 1 : 0;
  TGeom g;
  if (!tma_geom(dy, need_b1 + (has_y ? 1 : 0) + (has_norm ? 1 : 0), pad, &g)) return 1;
  const int es = elem_size(dy->dtype);
  NormTParams p;
  memset(&p, 0, sizeof(p));
  p.stats = stats;
  p.gamma = gamma;
  p.sums = sums;
  p.has_dy2 = has_dy2;
  p.has_y = has_y;
  p.has_x = has_norm;
  p.has_dres = has_dres;
  p.mode = a->mode;
  p.H = dy->h;
  p.W = dy->w;
  p.C = dy->c;
  p.N = dy->n;
  p.nv = g.slabb / 16;
  p.cblocks = dy->c * es / g.slabb;
  p.items = p.N * p.cblocks;
  p.cs = g.cs;
  p.R = g.R;
  p.dy_pad = pad;
  p.y_h = has_y ? y->halo : 0;
  p.off_b1 = g.dyb;
  p.off_y = g.dyb + need_b1 * g.dense;
  p.off_x = p.off_y + (has_y ? g.dense : 0);
  p.set_bytes = g.set_bytes;
  const int sc = g.slabb / es;
  const unsigned dense_tx = static_cast<unsigned>(g.R) * p.W * g.slabb;
  int rc;
  if (pad) {
    const int rows = g.R + (g.cs == 1 ? 2 : 1);
    rc = plane_map(&p.tm_dy, dy, sc, p.W + 2, rows);
    p.tx_bytes = static_cast<unsigned>(rows) * (p.W + 2) * g.slabb;
  } else {
    rc = plane_map(&p.tm_dy, dy, sc, p.W, g.R);
    p.tx_bytes = dense_tx;
  }
  if (rc != DTG_OK) return rc;
  if (has_dy2) {
    if ((rc = plane_map(&p.tm_dy2, dy2, sc, p.W, g.R)) != DTG_OK) return rc;
    p.tx_bytes += dense_tx;
  }
  if (has_y) {
    if ((rc = plane_map(&p.tm_y, y, sc, p.W, g.R)) != DTG_OK) return rc;
    p.tx_bytes += dense_tx;
  }
  if (has_norm) {
    if ((rc = plane_map(&p.tm_x, x, sc, p.W, g.R)) != DTG_OK) return rc;
    p.tx_bytes += dense_tx;
  }
  if ((rc = plane_map(&p.tm_dx, dx, sc, p.W, g.R)) != DTG_OK) return rc;
  if (has_dres && (rc = plane_map(&p.tm_dres, d_res, sc, p.W, g.R)) != DTG_OK) return rc;
  const size_t smem = static_cast<size_t>(g.set_bytes) + 1024;
#define DTG_TMA_BWD3(TT, AA)                                                                                          \
  do {                                                                                                                \
    if (g.ppt == 4) {                                                                                                 \
      if (pad) return launch_tma_norm(norm_bwd_tma_kernel<TT, AA, 4, true>, p, smem, stream);                         \
      return launch_tma_norm(norm_bwd_tma_kernel<TT, AA, 4, false>, p, smem, stream);                                 \
    }                                                                                                                 \
    if (pad) return launch_tma_norm(norm_bwd_tma_kernel<TT, AA, 8, true>, p, smem, stream);                           \
    return launch_tma_norm(norm_bwd_tma_kernel<TT, AA, 8, false>, p, smem, stream);                                   \
  } while (0)
#define DTG_TMA_BWD(TT)                                       \
  do {                                                        \
    if (a->act == DTG_ACT_RELU) DTG_TMA_BWD3(TT, DTG_ACT_RELU);   \
    if (a->act == DTG_ACT_LRELU) DTG_TMA_BWD3(TT, DTG_ACT_LRELU); \
    DTG_TMA_BWD3(TT, DTG_ACT_NONE);                           \
  } while (0)
  if (dy->dtype == DTG_BF16) DTG_TMA_BWD(__nv_bfloat16);
  DTG_TMA_BWD(float);
#undef DTG_TMA_BWD
#undef DTG_TMA_BWD3
}

int try_norm_fwd_tma(const dtg_norm_args* a, const dtg_plane* x, const dtg_plane* residual, const float* gamma,
                     const float* beta, float* stats, const dtg_plane* out, cudaStream_t stream) {
  if (norm_impl() != 1 || a->phase != 0 || (a->mode != DTG_NORM_INSTANCE && a->mode != DTG_NORM_COND_INSTANCE)) return 1;
  if (a->act == DTG_ACT_TANH) return 1;
  const bool has_res = residual && residual->ptr;
  if (x->halo != 0 || x->h * x->w < 2 || x->n > 65535) return 1;
  TGeom g;
  if (!tma_geom(x, has_res ? 1 : 0, false, &g)) return 1;       // buffers: x (y in place) | residual
  if (out->halo > 0 && g.cs * g.R != x->h) return 1;       // a ragged last box would spill into the output halo
  const int es = elem_size(x->dtype);
  NormFParams p;
  memset(&p, 0, sizeof(p));
  p.xptr = x->ptr;
  p.gamma = gamma;
  p.beta = beta;
  p.stats = stats;
  p.out = *out;
  p.has_res = has_res;
  p.res_h = has_res ? residual->halo : 0;
  p.mode = a->mode;
  p.act = a->act;
  p.eps = a->eps;
  p.H = x->h;
  p.W = x->w;
  p.C = x->c;
  p.N = x->n;
  p.nv = g.slabb / 16;
  p.cblocks = x->c * es / g.slabb;
  p.items = p.N * p.cblocks;
  p.cs = g.cs;
  p.R = g.R;
  p.off_res = g.dense;
  p.set_bytes = g.set_bytes;
  const int sc = g.slabb / es;
  p.tx_bytes = static_cast<unsigned>(g.R) * p.W * g.slabb * (has_res ? 2u : 1u);
  int rc;
  if ((rc = plane_map(&p.tm_x, x, sc, p.W, g.R)) != DTG_OK) return rc;
  if (has_res && (rc = plane_map(&p.tm_res, residual, sc, p.W, g.R)) != DTG_OK) return rc;
  if ((rc = plane_map(&p.tm_out, out, sc, p.W, g.R)) != DTG_OK) return rc;
  const size_t smem = static_cast<size_t>(g.set_bytes) + 1024;
  if (x->dtype == DTG_BF16) {
    if (g.ppt == 4) return launch_tma_norm(norm_fwd_tma_kernel<__nv_bfloat16, 4>, p, smem, stream);
    return launch_tma_norm(norm_fwd_tma_kernel<__nv_bfloat16, 8>, p, smem, stream);
  }
  if (g.ppt == 4) return launch_tma_norm(norm_fwd_tma_kernel<float, 4>, p, smem, stream);
  return launch_tma_norm(norm_fwd_tma_kernel<float, 8>, p, smem, stream);
}

}  // namespace dtg
