// TMA-staged (conditional) instance normalisation backward: persistent clusters, double-buffered slabs.
//
// A "slab" is one sample x one channel block (128 / 64 / 32 bytes of channels) x all pixels; a cluster of CS CTAs
// owns a slab at a time and splits its rows.  Every input of the CTA's row range (dy, dy2, y, x) is brought into
// shared memory by ONE TMA box each (cp.async.bulk.tensor, mbarrier complete_tx), so a CTA keeps a whole work item in
// flight without holding registers: the loads of item k+1 are issued before item k is reduced, which hides the
// block reduction, the cluster barrier and the arithmetic behind the HBM stream (the register-resident kernels in
// norm_fused.cu serialise load -> barrier -> store per CTA and reach a third of the HBM peak).  Pass 1 reduces
// (sum g, sum g*xhat) from shared memory (warp shuffles -> per-warp partials -> fixed-order block sum -> pushed into
// every peer's shared memory through DSMEM -> summed in rank order: deterministic, no atomics); pass 2 re-reads the
// still-resident inputs, writes dx (and the residual-branch gradient) in place and they leave through TMA stores.
// HBM traffic = one read of every input + one write of every output.
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

constexpr int kTN = 512;            // threads per CTA
constexpr int kTSlab = 64;          // max channels per slab
constexpr int kTMaxCluster = 8;
constexpr int kTSmemBudget = 200 * 1024;

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

// Fixed-order reduction of per-thread (s1[V], s2[V]) over the pixel lanes of the CTA and the CTAs of the cluster:
// xor tree over the lanes of a warp that own the same channel vector (nv apart) -> per-warp partials -> block sum in
// warp order -> pushed into slot `rank` of every peer's allpart[buf] through DSMEM -> after ONE cluster barrier every
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
      a += wred[w][tid * 2];
      b += wred[w][tid * 2 + 1];
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
      A += allpart[r][tid * 2];
      B += allpart[r][tid * 2 + 1];
    }
  }
  return make_float2(A, B);
}

template <typename T, int ACT>
__global__ void __launch_bounds__(kTN, 1) norm_bwd_tma_kernel(const __grid_constant__ NormTParams p) {
  constexpr int V = Vec<T>::N;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ float wred[kTN / 32][kTSlab * 2];
  __shared__ float allpart[2][kTMaxCluster][kTSlab * 2];
  __shared__ float4 kco[kTSlab];
  __shared__ __align__(8) uint64_t bar_full[2];
  pdl_trigger();
  cg::cluster_group cl = cg::this_cluster();
  const int cs = p.cs;
  const int rank = cs > 1 ? static_cast<int>(cl.block_rank()) : 0;
  const int cluster_id = blockIdx.x / cs;
  const int tid = threadIdx.x;
  const int nv = p.nv, slab_ch = nv * V, slabb = nv * 16;
  const int v = tid % nv, lane = tid / nv, lanes = kTN / nv;
  const int r0 = rank * p.R;
  const int npix = p.R * p.W;
  const bool has_norm = p.has_x != 0;
  const int dy_w = p.W + 2;                       // pitch (pixels) of the padded dy box
  const int dy_row0 = rank == 0 ? 1 : 0;          // local box row of the CTA's first interior row

  if (tid == 0) {
    tma_prefetch_desc(&p.tm_dy);
    tma_prefetch_desc(&p.tm_dx);
    if (p.has_dy2) tma_prefetch_desc(&p.tm_dy2);
    if (p.has_y) tma_prefetch_desc(&p.tm_y);
    if (p.has_x) tma_prefetch_desc(&p.tm_x);
    if (p.has_dres) tma_prefetch_desc(&p.tm_dres);
    mbar_init(&bar_full[0], 1);
    mbar_init(&bar_full[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  pdl_wait();

  auto issue = [&](int item, int set) {     // thread 0: all TMA loads of one work item into buffer set `set`
    const int n = item / p.cblocks, c0 = (item - n * p.cblocks) * slab_ch;
    uint8_t* sb = smem + set * p.set_bytes;
    uint64_t* bar = &bar_full[set];
    mbar_expect_tx(bar, p.tx_bytes);
    if (p.dy_pad)
      tma_load_4d(sb, &p.tm_dy, bar, c0, 0, rank == 0 ? 0 : r0 + 1, n);
    else
      tma_load_4d(sb, &p.tm_dy, bar, c0, 0, r0, n);
    if (p.has_dy2) tma_load_4d(sb + p.off_b1, &p.tm_dy2, bar, c0, 0, r0, n);
    if (p.has_y) tma_load_4d(sb + p.off_y, &p.tm_y, bar, c0, p.y_h, r0 + p.y_h, n);
    if (p.has_x) tma_load_4d(sb + p.off_x, &p.tm_x, bar, c0, 0, r0, n);
  };
  if (tid == 0) {
    if (cluster_id < p.items) issue(cluster_id, 0);
    if (cluster_id + p.nclusters < p.items) issue(cluster_id + p.nclusters, 1);
  }

  // g = (fold(dy) + dy2) * act'(y) of local pixel pl (row ly, column lx) from the resident buffers
  auto load_g = [&](const uint8_t* sb, int pl, int ly, int lx, float (&g)[V]) {
    if (p.dy_pad) {
      const int y = r0 + ly;
      const int cell = (ly + dy_row0) * dy_w + lx + 1;
      Vec<T>::load(sb + (cell * nv + v) * 16, g);
      // reflection-pad(1) backward: row 1 also receives the halo row above row 0, row H-2 the halo row below row H-1
      const int dr = y == 1 ? -2 * dy_w : (y == p.H - 2 ? 2 * dy_w : 0);
      const int dc = lx == 1 ? -2 : (lx == p.W - 2 ? 2 : 0);
      if (dr != 0) {
        float t[V];
        Vec<T>::load(sb + ((cell + dr) * nv + v) * 16, t);
#pragma unroll
        for (int i = 0; i < V; ++i) g[i] += t[i];
      }
      if (dc != 0) {
        float t[V];
        Vec<T>::load(sb + ((cell + dc) * nv + v) * 16, t);
#pragma unroll
        for (int i = 0; i < V; ++i) g[i] += t[i];
        if (dr != 0) {
          Vec<T>::load(sb + ((cell + dr + dc) * nv + v) * 16, t);
#pragma unroll
          for (int i = 0; i < V; ++i) g[i] += t[i];
        }
      }
    } else {
      Vec<T>::load(sb + (pl * nv + v) * 16, g);
    }
    if (p.has_dy2) {
      float t[V];
      Vec<T>::load(sb + p.off_b1 + (pl * nv + v) * 16, t);
#pragma unroll
      for (int i = 0; i < V; ++i) g[i] += t[i];
    }
    if (ACT != DTG_ACT_NONE) {
      float t[V];
      Vec<T>::load(sb + p.off_y + (pl * nv + v) * 16, t);
#pragma unroll
      for (int i = 0; i < V; ++i) g[i] = t[i] > 0.f ? g[i] : (ACT == DTG_ACT_LRELU ? 0.2f * g[i] : 0.f);
    }
  };

  int k = 0;
  for (int item = cluster_id; item < p.items; item += p.nclusters, ++k) {
    const int set = k & 1;
    const uint32_t par = static_cast<uint32_t>(k >> 1) & 1u;
    const int n = item / p.cblocks, c0 = (item - n * p.cblocks) * slab_ch;
    uint8_t* sb = smem + set * p.set_bytes;
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
    mbar_wait(&bar_full[set], par);

    // ---- pass 1: A = sum g, B = sum g * xhat over the CTA's rows
    for (int pl = lane; pl < npix; pl += lanes) {
      const int ly = pl / p.W, lx = pl - ly * p.W;
      if (r0 + ly >= p.H) break;
      float g[V];
      load_g(sb, pl, ly, lx, g);
      if (has_norm) {
        float f[V];
        Vec<T>::load(sb + p.off_x + (pl * nv + v) * 16, f);
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
      const float m = static_cast<float>(p.H * p.W);
      const float d = p.mode == DTG_NORM_COND_INSTANCE ? m - 1.f : m;
      if (has_norm) {
        const float ga = p.mode == DTG_NORM_COND_INSTANCE ? p.gamma[nc] : p.gamma[ch];
        kco[tid] = make_float4(p.stats[nc * 2 + 1] * ga, A / m, B / d, 0.f);
      } else {
        kco[tid] = make_float4(1.f, 0.f, 0.f, 0.f);
      }
      if (rank == 0) {
        p.sums[nc * 2] = A;
        p.sums[nc * 2 + 1] = B;
      }
    }
    __syncthreads();

    // ---- pass 2: dx = k0 * (g - kA - xhat * kB) in place of x, d_res = g in place of dy2
    float k0[V], kA[V], kB[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float4 kk = kco[v * V + i];
      k0[i] = kk.x;
      kA[i] = kk.y;
      kB[i] = kk.z;
    }
    uint8_t* outb = sb + (has_norm ? p.off_x : p.off_y);
    for (int pl = lane; pl < npix; pl += lanes) {
      const int ly = pl / p.W, lx = pl - ly * p.W;
      if (r0 + ly >= p.H) break;
      float g[V];
      load_g(sb, pl, ly, lx, g);
      if (p.has_dres) Vec<T>::store(sb + p.off_b1 + (pl * nv + v) * 16, g);
      if (has_norm) {
        float f[V];
        Vec<T>::load(sb + p.off_x + (pl * nv + v) * 16, f);
#pragma unroll
        for (int i = 0; i < V; ++i) g[i] = k0[i] * (g[i] - kA[i] - ((f[i] - mean[i]) * rstd[i]) * kB[i]);
      }
      Vec<T>::store(outb + (pl * nv + v) * 16, g);
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tma_store_4d(&p.tm_dx, outb, c0, 0, r0, n);
      if (p.has_dres) tma_store_4d(&p.tm_dres, sb + p.off_b1, c0, 0, r0, n);
      bulk_commit();
      const int next = item + 2 * p.nclusters;
      if (next < p.items) {
        bulk_wait_read<0>();          // the stores have finished reading this buffer set
        issue(next, set);
      }
    }
  }
  if (tid == 0) bulk_wait_read<0>();
}

// ---------------------------------------------------------------------------------------------------------------
// forward: stats (shifted sums, K = first pixel of the sample) -> y = act(x*a + b (+ residual)), same pipeline.
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

template <typename T>
__global__ void __launch_bounds__(kTN, 1) norm_fwd_tma_kernel(const __grid_constant__ NormFParams p) {
  constexpr int V = Vec<T>::N;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ float wred[kTN / 32][kTSlab * 2];
  __shared__ float allpart[2][kTMaxCluster][kTSlab * 2];
  __shared__ float2 coef[kTSlab];
  __shared__ __align__(8) uint64_t bar_full[2];
  pdl_trigger();
  cg::cluster_group cl = cg::this_cluster();
  const int cs = p.cs;
  const int rank = cs > 1 ? static_cast<int>(cl.block_rank()) : 0;
  const int cluster_id = blockIdx.x / cs;
  const int tid = threadIdx.x;
  const int nv = p.nv, slab_ch = nv * V;
  const int v = tid % nv, lane = tid / nv, lanes = kTN / nv;
  const int r0 = rank * p.R;
  const int npix = p.R * p.W;

  if (tid == 0) {
    tma_prefetch_desc(&p.tm_x);
    tma_prefetch_desc(&p.tm_out);
    if (p.has_res) tma_prefetch_desc(&p.tm_res);
    mbar_init(&bar_full[0], 1);
    mbar_init(&bar_full[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  pdl_wait();

  auto issue = [&](int item, int set) {
    const int n = item / p.cblocks, c0 = (item - n * p.cblocks) * slab_ch;
    uint8_t* sb = smem + set * p.set_bytes;
    uint64_t* bar = &bar_full[set];
    mbar_expect_tx(bar, p.tx_bytes);
    tma_load_4d(sb, &p.tm_x, bar, c0, 0, r0, n);
    if (p.has_res) tma_load_4d(sb + p.off_res, &p.tm_res, bar, c0, p.res_h, r0 + p.res_h, n);
  };
  if (tid == 0) {
    if (cluster_id < p.items) issue(cluster_id, 0);
    if (cluster_id + p.nclusters < p.items) issue(cluster_id + p.nclusters, 1);
  }

  int k = 0;
  for (int item = cluster_id; item < p.items; item += p.nclusters, ++k) {
    const int set = k & 1;
    const uint32_t par = static_cast<uint32_t>(k >> 1) & 1u;
    const int n = item / p.cblocks, c0 = (item - n * p.cblocks) * slab_ch;
    uint8_t* sb = smem + set * p.set_bytes;
    // shift K = first pixel of the sample (identical in every CTA of the cluster; an L2 hit)
    float K[V], s1[V], s2[V];
    Vec<T>::load(reinterpret_cast<const uint8_t*>(p.xptr) +
                     (static_cast<size_t>(n) * p.H * p.W * p.C + c0 + v * V) * sizeof(T), K);
#pragma unroll
    for (int i = 0; i < V; ++i) s1[i] = s2[i] = 0.f;
    mbar_wait(&bar_full[set], par);
    for (int pl = lane; pl < npix; pl += lanes) {
      if (r0 + pl / p.W >= p.H) break;
      float f[V];
      Vec<T>::load(sb + (pl * nv + v) * 16, f);
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
    for (int pl = lane; pl < npix; pl += lanes) {
      const int ly = pl / p.W, lx = pl - ly * p.W;
      const int py = r0 + ly;
      if (py >= p.H) break;
      float f[V];
      Vec<T>::load(sb + (pl * nv + v) * 16, f);
#pragma unroll
      for (int i = 0; i < V; ++i) f[i] = f[i] * ca[i] + cb[i];
      if (p.has_res) {
        float t[V];
        Vec<T>::load(sb + p.off_res + (pl * nv + v) * 16, t);
#pragma unroll
        for (int i = 0; i < V; ++i) f[i] += t[i];
      }
#pragma unroll
      for (int i = 0; i < V; ++i) f[i] = apply_act(f[i], p.act);
      Vec<T>::store(sb + (pl * nv + v) * 16, f);
      if (p.out.halo > 0) {
        int hts[3], wts[3];
        const int nh = reflect_targets(py, p.out.h, p.out.halo, hts), nw = reflect_targets(lx, p.out.w, p.out.halo, wts);
        if (nh * nw > 1) {
          uint8_t* ob = reinterpret_cast<uint8_t*>(p.out.ptr);
          for (int a = 0; a < nh; ++a)
            for (int q = 0; q < nw; ++q)
              if (a + q > 0)
                Vec<T>::store(ob + (plane_pix(p.out, n, hts[a], wts[q]) * p.out.c + c0 + v * V) * sizeof(T), f);
        }
      }
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tma_store_4d(&p.tm_out, sb, c0, p.out.halo, r0 + p.out.halo, n);
      bulk_commit();
      const int next = item + 2 * p.nclusters;
      if (next < p.items) {
        bulk_wait_read<0>();
        issue(next, set);
      }
    }
  }
  if (tid == 0) bulk_wait_read<0>();
}

struct TGeom {
  int slabb, cs, R, dense, dyb, set_bytes, need_b1;
};

static bool tma_geom(const dtg_plane* dy, bool has_dy2, bool has_y, bool has_x, bool has_dres, TGeom* g) {
  const int es = elem_size(dy->dtype);
  const int rowb = dy->c * es;
  const int H = dy->h, W = dy->w, N = dy->n;
  const bool pad = dy->halo == 1;
  if (dy->halo > 1 || (pad && (H < 4 || W < 4))) return false;
  if (W > 254 || H < 1) return false;
  const int need_b1 = (has_dy2 || has_dres) ? 1 : 0;
  // widest slab (DRAM-friendly rows) whose double-buffered footprint fits; smallest cluster that fits and still gives
  // the GPU >= 148 CTAs (else the largest feasible cluster)
  for (int slabb = 128; slabb >= 32; slabb >>= 1) {
    if (rowb % slabb != 0) continue;
    const int items = N * (rowb / slabb);
    bool found = false;
    for (int cs = 1; cs <= kTMaxCluster; cs *= 2) {
      const int R = (H + cs - 1) / cs;
      if ((cs - 1) * R >= H) break;                         // a CTA without rows
      if (pad && (H % cs != 0 || R < 2)) continue;
      if (R + 2 > 256) continue;
      const int dense = (R * W * slabb + 127) & ~127;
      const int dyb = pad ? (((R + (cs == 1 ? 2 : 1)) * (W + 2) * slabb + 127) & ~127) : dense;
      const int set = dyb + (need_b1 + (has_y ? 1 : 0) + (has_x ? 1 : 0)) * dense;
      if (2 * set > kTSmemBudget) continue;
      g->slabb = slabb;
      g->cs = cs;
      g->R = R;
      g->dense = dense;
      g->dyb = dyb;
      g->set_bytes = (set + 1023) & ~1023;
      g->need_b1 = need_b1;
      found = true;
      if (items * cs >= 148) break;
    }
    if (found) return true;
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
      DTG_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 20 * 1024));
      cfg.gridDim = dim3(p.cs * 148, 1, 1);
      cfg.numAttrs = 1;
      DTG_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&ncl, kernel, &cfg));
      if (ncl < 1) ncl = 1;
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
  static const bool disabled = getenv("DTG_NO_TMA_NORM") != nullptr;
  if (disabled || a->phase != 0 || a->mode == DTG_NORM_BATCH) return 1;
  const bool has_norm = a->mode != DTG_NORM_NONE;
  const bool has_dy2 = dy2 && dy2->ptr, has_y = a->act != DTG_ACT_NONE, has_dres = d_res && d_res->ptr;
  if (!has_norm && !has_y) return 1;
  if (has_y && !(y && y->ptr)) return 1;
  if (dy->n > 65535 || dx->halo != 0 || (has_dres && d_res->halo != 0) || (has_dy2 && dy2->halo != 0)) return 1;
  TGeom g;
  if (!tma_geom(dy, has_dy2, has_y, has_norm, has_dres, &g)) return 1;
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
  p.dy_pad = dy->halo == 1;
  p.y_h = has_y ? y->halo : 0;
  p.off_b1 = g.dyb;
  p.off_y = g.dyb + g.need_b1 * g.dense;
  p.off_x = p.off_y + (has_y ? g.dense : 0);
  p.set_bytes = g.set_bytes;
  const int sc = g.slabb / es;
  const unsigned dense_tx = static_cast<unsigned>(g.R) * p.W * g.slabb;
  int rc;
  if (p.dy_pad) {
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
  const size_t smem = 2 * static_cast<size_t>(g.set_bytes) + 1024;
#define DTG_TMA_BWD(TT)                                                                                     \
  do {                                                                                                      \
    if (a->act == DTG_ACT_RELU) return launch_tma_norm(norm_bwd_tma_kernel<TT, DTG_ACT_RELU>, p, smem, stream);   \
    if (a->act == DTG_ACT_LRELU) return launch_tma_norm(norm_bwd_tma_kernel<TT, DTG_ACT_LRELU>, p, smem, stream); \
    return launch_tma_norm(norm_bwd_tma_kernel<TT, DTG_ACT_NONE>, p, smem, stream);                               \
  } while (0)
  if (dy->dtype == DTG_BF16) DTG_TMA_BWD(__nv_bfloat16);
  DTG_TMA_BWD(float);
#undef DTG_TMA_BWD
}

int try_norm_fwd_tma(const dtg_norm_args* a, const dtg_plane* x, const dtg_plane* residual, const float* gamma,
                     const float* beta, float* stats, const dtg_plane* out, cudaStream_t stream) {
  static const bool disabled = getenv("DTG_NO_TMA_NORM") != nullptr;
  if (disabled || a->phase != 0 || (a->mode != DTG_NORM_INSTANCE && a->mode != DTG_NORM_COND_INSTANCE)) return 1;
  const bool has_res = residual && residual->ptr;
  if (x->halo != 0 || x->h * x->w < 2 || x->n > 65535) return 1;
  TGeom g;
  // buffers per set: x (dense, y in place) + residual
  if (!tma_geom(x, has_res, false, false, false, &g)) return 1;
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
  const size_t smem = 2 * static_cast<size_t>(g.set_bytes) + 1024;
  if (x->dtype == DTG_BF16) return launch_tma_norm(norm_fwd_tma_kernel<__nv_bfloat16>, p, smem, stream);
  return launch_tma_norm(norm_fwd_tma_kernel<float>, p, smem, stream);
}

}  // namespace dtg
