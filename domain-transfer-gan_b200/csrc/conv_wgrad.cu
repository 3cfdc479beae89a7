// Weight gradient as a tcgen05 GEMM whose K dimension is the pixel axis.
//
//   dW[a][b][tap] += sum_pixels P[pix][a] * Q[pix*stride + tap - pad][b]
//
// Both operands are read straight from the NHWC planes by TMA: a tile of 64 pixels x 128 bytes of
// channels lands in smem as 64 rows of 128 B (SWIZZLE_128B) = an MN-major UMMA operand (pixels are
// the K rows).  M = 128 P-channels (two/four 128-byte channel blocks, LBO apart), N = Q channels,
// one TMEM accumulator [128 x N] per filter tap; a CTA owns a group of taps (<= 512 TMEM columns),
// one 128-channel M block and a contiguous range of pixel tiles (split-K).  Partial tiles go to a
// workspace with plain coalesced stores; wgrad_reduce_kernel sums the splits in a fixed order and
// accumulates into the fp32 PyTorch-layout gradient (deterministic, no atomics).
#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace dtg {

constexpr int kWMaxTaps = 64;
constexpr int kKP = 64;                 // pixels (K rows) per stage (patch mode may use 128: WgradParams::kp)
constexpr int kBlkBytes = kKP * 128;    // one 128-byte-wide channel block of a 64-pixel tile: 8 KB
constexpr int kOffTab = 128;            // patch-mode window-offset table entries (taps per group x k-steps)
constexpr int kWThreads = 192;

struct WgradParams {
  CUtensorMap tmP;
  CUtensorMap tmQ[4];
  int bw, bh, bn;
  int tiles_w, tiles_h, tiles_n;
  int tiles_per_split;
  int ntaps, tpg;
  short tap_dh[kWMaxTaps], tap_dw[kWMaxTaps];
  unsigned char tap_map[kWMaxTaps];
  int nblkA, nblkB;
  int n_umma;
  int mtot;  // mblocks * 128
  int stages, tmem_cols;
  // patch mode (stride 1): the q operand of a tap group (whole filter rows) is ONE haloed patch per channel block
  // per stage; every tap is a row-shifted window of it (UMMA descriptors with shifted start addresses)
  int patch;
  int pw, ph;                 // patch extents (pixels)
  int q_org_w, q_org_h;       // patch origin = tile origin + org (+ group * rows_per_group in h)
  int rows_per_group;
  int q_blk_bytes;            // pw * ph * 128
  int stage_bytes;
  unsigned short b_off[kOffTab];     // [local tap][k-step] window start inside the patch, 16-byte units
  int kp;                            // pixels per stage (64 or 128)
  int blk_bytes;                     // kp * 128: one channel block of the p tile
  int a_load_blocks;                 // p channel blocks actually loaded; the others read a shared all-zero block
  int zero_off;                      // byte offset of the zero block from the smem base
  float* ws;
  int rows_valid, cols_valid;        // accumulator rows / columns that hold data (the rest is zero padding: never stored)
  int dbg;                           // DTG_WGRAD_DBG experiments: 1 = no MMAs, 2 = no TMA loads, 4 = no epilogue stores
};

template <bool TF32>
__global__ void __launch_bounds__(kWThreads, 1) wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  pdl_trigger();      // the next kernel may start its prologue; it still waits for this grid before touching memory
  constexpr int CH = TF32 ? 32 : 64;   // channels per 128-byte row
  constexpr int UK = TF32 ? 8 : 16;    // pixels consumed per MMA
  const int split = blockIdx.x, group = blockIdx.y, mblock = blockIdx.z;
  const int tap0 = group * p.tpg;
  const int ntl = min(p.tpg, p.ntaps - tap0);
  const int a_bytes = p.a_load_blocks * p.blk_bytes;
  const int stage_bytes = p.stage_bytes;
  const int S = p.stages;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + S * stage_bytes);
  uint64_t* bar_empty = bar_full + S;
  uint64_t* bar_done = bar_empty + S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmP);
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.tmQ[i]);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < S; ++i) {
        mbar_init(&bar_full[i], 1);
        mbar_init(&bar_empty[i], 1);
      }
      mbar_init(bar_done, 1);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, p.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();         // barriers / TMEM / descriptor prefetch above overlap the previous kernel's tail

  const int T = p.tiles_w * p.tiles_h * p.tiles_n;
  const int t_begin = split * p.tiles_per_split;
  const int t_end = min(T, t_begin + p.tiles_per_split);
  const uint32_t tx_bytes = p.patch ? static_cast<uint32_t>(a_bytes + p.nblkB * p.q_blk_bytes)
                                    : static_cast<uint32_t>(a_bytes + ntl * p.nblkB * p.blk_bytes);
  if (p.a_load_blocks < p.nblkA) {     // all-zero p block for the channel blocks that hold no data (pa <= 64)
    for (int i = threadIdx.x * 16; i < p.blk_bytes; i += kWThreads * 16)
      *reinterpret_cast<uint4*>(smem + p.zero_off + i) = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
  }

  if (warp == 0) {
    {   // warp-uniform loop; only the TMA / mbarrier instructions are predicated on elect.sync
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = t_begin; tile < t_end; ++tile) {
        int r = tile;
        const int tw = r % p.tiles_w;
        r /= p.tiles_w;
        const int th = r % p.tiles_h;
        const int tn = r / p.tiles_h;
        const int a0 = th * p.bh, b0 = tw * p.bw, n0 = tn * p.bn;
        mbar_wait(&bar_empty[stage], phase ^ 1);
        uint8_t* s = smem + stage * stage_bytes;
        if (p.dbg & 2) {
          if (elect_one()) mbar_arrive(&bar_full[stage]);
          __syncwarp();
          if (++stage == S) {
            stage = 0;
            phase ^= 1;
          }
          continue;
        }
        if (elect_one()) mbar_expect_tx(&bar_full[stage], tx_bytes);
        __syncwarp();
        for (int blk = 0; blk < p.a_load_blocks; ++blk)
          if (elect_one()) tma_load_4d(s + blk * p.blk_bytes, &p.tmP, &bar_full[stage], mblock * 128 + blk * CH, b0, a0, n0);
        if (p.patch) {
          for (int blk = 0; blk < p.nblkB; ++blk)
            if (elect_one())
              tma_load_4d(s + a_bytes + blk * p.q_blk_bytes, &p.tmQ[0], &bar_full[stage], blk * CH, b0 + p.q_org_w,
                          a0 + p.q_org_h + group * p.rows_per_group, n0);
        } else {
          for (int tl = 0; tl < ntl; ++tl) {
            const int t = tap0 + tl;
            for (int blk = 0; blk < p.nblkB; ++blk)
              if (elect_one())
                tma_load_4d(s + a_bytes + (tl * p.nblkB + blk) * p.blk_bytes, &p.tmQ[p.tap_map[t]], &bar_full[stage],
                            blk * CH, b0 + p.tap_dw[t], a0 + p.tap_dh[t], n0);
          }
        }
        __syncwarp();
        if (++stage == S) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc(TF32 ? 2u : 1u, 1u, 1u, 128, p.n_umma);
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = t_begin; tile < t_end; ++tile) {
      mbar_wait(&bar_full[stage], phase);
      tc_fence_after();
      {
        const uint32_t sa = smem_u32(smem + stage * stage_bytes);
        // second p channel block: either loaded right behind the first or the shared zero block (LBO reaches it)
        const uint32_t a_lbo = p.a_load_blocks < p.nblkA ? smem_u32(smem + p.zero_off) - sa : static_cast<uint32_t>(p.blk_bytes);
        const uint64_t ad0 = umma_desc_sw128(sa, a_lbo, TF32 ? 512 : 1024, TF32 ? 1 : 2);
        const int nk = (p.dbg & 1) ? 0 : p.kp / UK;
        if (p.patch) {
          const uint64_t bq0 = umma_desc_sw128(sa + a_bytes, p.q_blk_bytes, TF32 ? 512 : 1024, TF32 ? 1 : 2);
          if (elect_one()) {
            const uint32_t acc = tile > t_begin ? 1u : 0u;
            for (int tl = 0; tl < ntl; ++tl) {
              const uint32_t d = tmem_base + tl * p.n_umma;
#pragma unroll 4
              for (int j = 0; j < nk; ++j)
                tc_mma<TF32>(d, ad0 + j * (UK * 128 / 16), bq0 + p.b_off[tl * nk + j], idesc, j > 0 ? 1u : acc);
            }
          }
        } else {
          for (int tl = 0; tl < ntl; ++tl) {
            const uint32_t sb = sa + a_bytes + tl * p.nblkB * p.blk_bytes;
            const uint64_t bd0 = umma_desc_sw128(sb, p.blk_bytes, TF32 ? 512 : 1024, TF32 ? 1 : 2);
            const uint32_t d = tmem_base + tl * p.n_umma;
            if (elect_one()) {
#pragma unroll 4
              for (int j = 0; j < nk; ++j)
                tc_mma<TF32>(d, ad0 + j * (UK * 128 / 16), bd0 + j * (UK * 128 / 16), idesc, (tile > t_begin || j > 0) ? 1u : 0u);
            }
          }
        }
        if (elect_one()) tc_commit(&bar_empty[stage]);
      }
      __syncwarp();
      if (++stage == S) {
        stage = 0;
        phase ^= 1;
      }
    }
    if (elect_one()) tc_commit(bar_done);
    __syncwarp();
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    mbar_wait(bar_done, 0);
    tc_fence_after();
    const bool any = t_end > t_begin;
    // workspace layout [split][tap][column / 4][row][4]: the 32 lanes of a warp hold 32 consecutive rows, so every
    // float4 store instruction writes 512 contiguous bytes
    const bool rows_live = mblock * 128 + quad * 32 < p.rows_valid;      // warp-uniform
    for (int tl = 0; tl < (rows_live ? ntl : 0); ++tl) {
      const int t = tap0 + tl;
      float4* dst = reinterpret_cast<float4*>(p.ws) + (static_cast<size_t>(split) * p.ntaps + t) * (p.n_umma / 4) * p.mtot +
                    mblock * 128 + row;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + tl * p.n_umma;
      for (int c0 = 0; c0 < p.cols_valid; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 o = any ? make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                       __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]))
                         : make_float4(0.f, 0.f, 0.f, 0.f);
          if (!(p.dbg & 4)) dst[static_cast<size_t>(c0 / 4 + q) * p.mtot] = o;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// dw[(a*qb + b)*ntaps + t] += sum_s ws[(((s*ntaps + t)*(n_umma/4) + b/4)*mtot + a)*4 + b%4]
// fold = 1: accumulator column b' = j*fc + b holds filter column kw = j          (t = kh, KW real columns)
// fold = 2: accumulator row    a' = j*fc + a holds filter column kw = KW - 1 - j
// kRSub (default 4) threads per float4 of accumulator columns: lane q of the group loads splits q, q+kRSub, ... (a handful of
// independent 16-byte loads; the memory-level parallelism comes from the number of resident warps, ptxas sinks
// per-thread load batches behind the adds), sums them in split order, then the lane sums are combined by a fixed xor
// tree: deterministic, no shared memory.  A warp load instruction covers 32 / kRSub consecutive rows of kRSub splits.
template <int kRSub>
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float4* __restrict__ ws, float* __restrict__ dw, int splits,
                                                           int ntaps, int mtot, int n_umma, int pa, int qb, int fold, int KW,
                                                           int fc, int rows, int ncv4, int total, int dbg) {
  pdl_enter();
  const int gi = blockIdx.x * 256 + threadIdx.x;
  const int i = gi / kRSub, sub = gi % kRSub;
  const bool ok = i < total;
  const int ncol4 = n_umma / 4;
  const int row = ok ? i % rows : 0;
  const int c4 = ok ? (i / rows) % ncv4 : 0;
  const int t = ok ? i / (ncv4 * rows) : 0;
  const float4* src = ws + (static_cast<size_t>(t) * ncol4 + c4) * mtot + row;
  const size_t sstride = static_cast<size_t>(ntaps) * mtot * ncol4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ok) {
#pragma unroll 4
    for (int s = sub; s < splits; s += kRSub) {
      const float4 v = __ldcg(src + s * sstride);
      acc.x += v.x;
      acc.y += v.y;
      acc.z += v.z;
      acc.w += v.w;
    }
  }
#pragma unroll
  for (int off = 1; off < kRSub; off <<= 1) {
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, off);
    acc.w += __shfl_xor_sync(0xffffffffu, acc.w, off);
  }
  if (!ok || sub != 0 || (dbg & 8)) return;
  const float av[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int col = c4 * 4 + e;
    int a = row, b = col, j = 0;
    if (fold == 1) {
      j = col / fc;
      b = col % fc;
    } else if (fold == 2) {
      j = row / fc;
      a = row % fc;
    }
    if (a >= pa || b >= qb || j >= KW) continue;
    // every gradient element is owned by exactly ONE thread of this launch, so the accumulation into dw is a single
    // red.global.add per address: deterministic, and fire-and-forget (a load-add-store here costs a dependent L2 round
    // trip per element on a scattered, 36-byte-stride pattern: measured 17 of the kernel's 23 us)
    if (fold) {
      const int kw = fold == 2 ? KW - 1 - j : j;
      atomicAdd(&dw[((static_cast<size_t>(a) * qb + b) * ntaps + t) * KW + kw], av[e]);
    } else {
      atomicAdd(&dw[(static_cast<size_t>(a) * qb + b) * ntaps + t], av[e]);
    }
  }
}

struct WgradPlan {
  int bw, bh, bn, tiles_w, tiles_h, tiles_n, T;
  int ntaps, tpg, ngroups, mblocks, nblkA, nblkB, n_umma, stages, tmem_cols, splits, tiles_per_split;
  int patch, pw, ph, rows_per_group, stage_bytes, kp, a_load_blocks;
  size_t ws_bytes;
};

static int pow2ceil(int v) {
  int r = 1;
  while (r < v) r <<= 1;
  return r;
}

static int make_plan(const dtg_wgrad_args* a, const dtg_plane* pp, const dtg_plane* q, WgradPlan* pl) {
  const bool tf32 = pp->dtype == DTG_F32;
  const int CH = tf32 ? 32 : 64;
  DTG_REQUIRE(pp->dtype == q->dtype, "wgrad: dtype mismatch");
  DTG_REQUIRE(a->kh * a->kw <= kWMaxTaps, "wgrad: too many taps");
  DTG_REQUIRE(a->stride == 1 || a->stride == 2, "wgrad: stride");
  const int fold = a->fold;
  DTG_REQUIRE(fold >= 0 && fold <= 2, "wgrad: bad fold mode");
  // p (the output gradient) may carry a halo ring (ignored: the tensor map views the interior); fold 2 reads its own zero halo
  if (fold) {
    const int es_ = tf32 ? 4 : 2;
    DTG_REQUIRE(a->stride == 1 && a->kw <= 8, "wgrad fold: stride 1 and kw <= 8 only");
    if (fold == 1)
      DTG_REQUIRE(q->c * es_ == 16 && q->halo >= a->pad && q->halo >= a->kw - 1 - a->pad,
                  "wgrad fold 1: q must be a 16-byte-per-pixel plane with a materialised halo >= pad");
    else
      DTG_REQUIRE(pp->c * es_ == 16 && pp->halo >= a->kw - 1 - a->pad && pp->halo >= a->pad && q->halo == 0,
                  "wgrad fold 2: p must be a 16-byte-per-pixel plane with a zero halo >= pad, q without halo");
  }
  DTG_REQUIRE(pp->n == q->n, "wgrad: batch mismatch");
  DTG_REQUIRE((q->h + 2 * a->pad - a->kh) / a->stride + 1 == pp->h && (q->w + 2 * a->pad - a->kw) / a->stride + 1 == pp->w,
              "wgrad: p extent %dx%d inconsistent with q %dx%d k%d s%d p%d", pp->h, pp->w, q->h, q->w, a->kh, a->stride, a->pad);
  DTG_REQUIRE(q->halo == 0 || q->halo >= a->pad, "wgrad: q halo %d < pad %d", q->halo, a->pad);
  DTG_REQUIRE(a->pa <= pp->c && a->qb <= q->c, "wgrad: valid channels exceed plane channels");
  pl->ntaps = fold ? a->kh : a->kh * a->kw;
  const int pa_eff = fold == 2 ? CH : a->pa;     // folded operand: 8 kw slots x 16 bytes = one 128-byte channel block
  const int qb_eff = fold == 1 ? CH : a->qb;
  pl->kp = kKP;
  pl->bw = std::min(pow2ceil(pp->w), kKP);
  pl->bh = std::min(pow2ceil(pp->h), kKP / pl->bw);
  pl->bn = kKP / (pl->bw * pl->bh);
  pl->mblocks = (pa_eff + 127) / 128;
  pl->nblkA = 128 / CH;
  // p channel blocks that hold data in the LAST (or only) M block; with one M block and <= 64 channels the second
  // block is never loaded: the A descriptor's LBO points at a shared all-zero block instead
  pl->a_load_blocks = (pl->nblkA == 2 && pl->mblocks == 1 && pa_eff <= CH) ? 1 : pl->nblkA;
  pl->n_umma = (qb_eff + CH - 1) / CH * CH;
  DTG_REQUIRE(pl->n_umma <= 256, "wgrad: q channels %d > 256", a->qb);
  pl->nblkB = pl->n_umma / CH;
  const int tmem_max = 512 / pl->n_umma;
  const int smem_max = (100 * 1024 - pl->a_load_blocks * kBlkBytes) / (pl->nblkB * kBlkBytes);
  const int tpg_max = std::max(1, std::min(tmem_max, smem_max));
  pl->ngroups = (pl->ntaps + tpg_max - 1) / tpg_max;
  pl->tpg = (pl->ntaps + pl->ngroups - 1) / pl->ngroups;
  pl->ngroups = (pl->ntaps + pl->tpg - 1) / pl->tpg;
  int stage_bytes = (pl->a_load_blocks + pl->tpg * pl->nblkB) * kBlkBytes;
  // patch mode: groups are whole filter rows (kh_eff x kw_eff taps, folded layers have kw_eff = 1); 128-pixel stages
  // (32 x 4 tiles) when they fit, so that the haloed patch amortises over more pixels
  pl->patch = 0;
  {
    static const bool no_patch = getenv("DTG_NO_WGRAD_PATCH") != nullptr;
    const int kh_eff = a->kh, kw_eff = fold ? 1 : a->kw;
    const int UKp = tf32 ? 8 : 16;
    for (int kp = 128; kp >= 64 && !pl->patch && !no_patch; kp -= 64) {
      const int bw = std::min(pow2ceil(pp->w), kp == 128 ? 32 : kKP);
      const int bh = std::min(pow2ceil(pp->h), kp / bw);
      if (a->stride != 1 || bw * bh != kp || bw < UKp || kw_eff * pl->n_umma > 512) continue;
      for (int cand = kh_eff; cand >= 1; --cand) {
        if (kh_eff % cand != 0 || cand * kw_eff * pl->n_umma > 512 || cand * kw_eff * (kp / UKp) > kOffTab) continue;
        const int sb = pl->a_load_blocks * kp * 128 + pl->nblkB * (bw + kw_eff - 1) * (bh + cand - 1) * 128;
        if ((kp == 128 ? 3 : 2) * ((sb + 1023) & ~1023) > 180 * 1024) continue;
        pl->patch = 1;
        pl->kp = kp;
        pl->bw = bw;
        pl->bh = bh;
        pl->bn = 1;
        pl->rows_per_group = cand;
        pl->pw = bw + kw_eff - 1;
        pl->ph = bh + cand - 1;
        pl->tpg = cand * kw_eff;
        pl->ngroups = kh_eff / cand;
        stage_bytes = sb;
        break;
      }
    }
  }
  pl->tiles_w = (pp->w + pl->bw - 1) / pl->bw;
  pl->tiles_h = (pp->h + pl->bh - 1) / pl->bh;
  pl->tiles_n = (pp->n + pl->bn - 1) / pl->bn;
  pl->T = pl->tiles_w * pl->tiles_h * pl->tiles_n;
  stage_bytes = (stage_bytes + 1023) & ~1023;
  pl->stage_bytes = stage_bytes;
  pl->stages = std::max(2, std::min(6, std::min(180 * 1024, tensor_smem_budget() - 12 * 1024) / stage_bytes));
  int cols = 32;
  while (cols < pl->tpg * pl->n_umma) cols <<= 1;
  pl->tmem_cols = cols;
  int want = std::max(1, 148 / (pl->ngroups * pl->mblocks));
  want = std::min(want, pl->T);
  pl->tiles_per_split = (pl->T + want - 1) / want;
  pl->splits = (pl->T + pl->tiles_per_split - 1) / pl->tiles_per_split;
  pl->ws_bytes = static_cast<size_t>(pl->splits) * pl->ntaps * pl->mblocks * 128 * pl->n_umma * sizeof(float);
  return DTG_OK;
}

static int floordiv2w(int e) { return (e - (e & 1)) / 2; }

}  // namespace dtg

using namespace dtg;

extern "C" size_t dtg_conv_wgrad_workspace_bytes(const dtg_wgrad_args* a, const dtg_plane* p, const dtg_plane* q) {
  WgradPlan pl;
  if (!a || !p || !q || make_plan(a, p, q, &pl) != DTG_OK) return 0;
  return pl.ws_bytes;
}

extern "C" int dtg_conv_wgrad(const dtg_wgrad_args* a, const dtg_plane* pp, const dtg_plane* q, float* dw,
                              void* workspace, size_t workspace_bytes, void* stream_) {
  DTG_REQUIRE(a && pp && q && dw && pp->ptr && q->ptr, "dtg_conv_wgrad: null argument");
  WgradPlan pl;
  int rc = make_plan(a, pp, q, &pl);
  if (rc != DTG_OK) return rc;
  DTG_REQUIRE(workspace && workspace_bytes >= pl.ws_bytes, "dtg_conv_wgrad: workspace %zu < %zu bytes", workspace_bytes, pl.ws_bytes);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool tf32 = pp->dtype == DTG_F32;
  const int es = elem_size(pp->dtype);
  const int CH = 128 / es;

  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.bw = pl.bw; p.bh = pl.bh; p.bn = pl.bn;
  p.tiles_w = pl.tiles_w; p.tiles_h = pl.tiles_h; p.tiles_n = pl.tiles_n;
  p.tiles_per_split = pl.tiles_per_split;
  p.ntaps = pl.ntaps; p.tpg = pl.tpg;
  p.nblkA = pl.nblkA; p.nblkB = pl.nblkB; p.n_umma = pl.n_umma;
  p.mtot = pl.mblocks * 128;
  p.stages = pl.stages; p.tmem_cols = pl.tmem_cols;
  p.patch = pl.patch; p.pw = pl.pw; p.ph = pl.ph; p.rows_per_group = pl.rows_per_group;
  p.q_blk_bytes = pl.pw * pl.ph * 128;
  p.stage_bytes = pl.stage_bytes;
  p.kp = pl.kp;
  p.blk_bytes = pl.kp * 128;
  p.a_load_blocks = pl.a_load_blocks;
  p.zero_off = pl.stages * pl.stage_bytes + 1024;      // behind the stages and the barrier block
  p.ws = reinterpret_cast<float*>(workspace);
  p.rows_valid = a->fold == 2 ? a->kw * (16 / es) : a->pa;
  p.cols_valid = std::min(pl.n_umma, ((a->fold == 1 ? a->kw * (16 / es) : a->qb) + 15) / 16 * 16);
  {
    static const int dbgv = getenv("DTG_WGRAD_DBG") ? atoi(getenv("DTG_WGRAD_DBG")) : 0;
    p.dbg = dbgv;
  }

  const int s = a->stride, hl = q->halo;
  const int fold = a->fold;
  if (fold)
    for (int kh = 0; kh < a->kh; ++kh) {
      p.tap_dh[kh] = static_cast<short>(kh - a->pad + hl);
      p.tap_dw[kh] = static_cast<short>(fold == 1 ? hl - a->pad : 0);
      p.tap_map[kh] = 0;
    }
  for (int kh = 0; kh < a->kh && !fold; ++kh)
    for (int kw = 0; kw < a->kw; ++kw) {
      const int t = kh * a->kw + kw;
      const int eh = kh - a->pad + hl, ew = kw - a->pad + hl;
      if (s == 1) {
        p.tap_dh[t] = static_cast<short>(eh);
        p.tap_dw[t] = static_cast<short>(ew);
        p.tap_map[t] = 0;
      } else {
        p.tap_dh[t] = static_cast<short>(floordiv2w(eh));
        p.tap_dw[t] = static_cast<short>(floordiv2w(ew));
        p.tap_map[t] = static_cast<unsigned char>((eh & 1) * 2 + (ew & 1));
      }
    }
  if (pl.patch) {
    // tap (kh, kw) of group g = kh / r reads q at tile pixel + (tap_dh, tap_dw); patch origin = offsets of the group's
    // first tap; local tap order is kh-major, matching the global tap index t = kh * kw_eff + kw
    const int kw_eff = fold ? 1 : a->kw;
    const int UKp = tf32 ? 8 : 16, nk = pl.kp / UKp, segs = pl.bw / UKp;
    p.q_org_h = p.tap_dh[0];
    p.q_org_w = p.tap_dw[0];
    for (int tl = 0; tl < pl.tpg; ++tl) {
      const int dkh = tl / kw_eff, dkw = tl % kw_eff;
      for (int j = 0; j < nk; ++j) {
        const int row = j / segs, seg = j % segs;
        p.b_off[tl * nk + j] = static_cast<unsigned short>((((dkh + row) * pl.pw + dkw + seg * UKp) * 128) >> 4);
      }
    }
  }
  uint32_t box[4] = {static_cast<uint32_t>(CH), static_cast<uint32_t>(pl.bw), static_cast<uint32_t>(pl.bh),
                     static_cast<uint32_t>(pl.bn)};
  uint32_t qbox[4] = {static_cast<uint32_t>(CH), static_cast<uint32_t>(pl.patch ? pl.pw : pl.bw),
                      static_cast<uint32_t>(pl.patch ? pl.ph : pl.bh), static_cast<uint32_t>(pl.bn)};
  if (fold == 2) {
    // p folded: window of pixel x = padded pixels x + halo - (KW-1-pad) .. +7 (overlapping rows, zero halo)
    const int hp = pp->halo, Hp = pp->h + 2 * hp, Wp = pp->w + 2 * hp;
    const int sft = hp - (a->kw - 1 - a->pad);
    uint8_t* base = reinterpret_cast<uint8_t*>(pp->ptr) + (static_cast<size_t>(hp) * Wp + sft) * pp->c * es;
    uint64_t dims[4] = {static_cast<uint64_t>(CH), static_cast<uint64_t>(pp->w), static_cast<uint64_t>(pp->h),
                        static_cast<uint64_t>(pp->n)};
    uint64_t strides[3] = {static_cast<uint64_t>(pp->c) * es, static_cast<uint64_t>(Wp) * pp->c * es,
                           static_cast<uint64_t>(Hp) * Wp * pp->c * es};
    rc = encode_tiled(&p.tmP, pp->dtype, 4, base, dims, strides, box, tf32 ? 2 : 1);
    if (rc != DTG_OK) return rc;
  } else {
    const int hp = pp->halo, Hp = pp->h + 2 * hp, Wp = pp->w + 2 * hp;      // interior view of a haloed plane
    uint8_t* base = reinterpret_cast<uint8_t*>(pp->ptr) + (static_cast<size_t>(hp) * Wp + hp) * pp->c * es;
    uint64_t dims[4] = {static_cast<uint64_t>(pp->c), static_cast<uint64_t>(pp->w), static_cast<uint64_t>(pp->h),
                        static_cast<uint64_t>(pp->n)};
    uint64_t strides[3] = {static_cast<uint64_t>(pp->c) * es, static_cast<uint64_t>(Wp) * pp->c * es,
                           static_cast<uint64_t>(Hp) * Wp * pp->c * es};
    rc = encode_tiled(&p.tmP, pp->dtype, 4, base, dims, strides, box, tf32 ? 2 : 1);
    if (rc != DTG_OK) return rc;
  }
  const int Hb = q->h + 2 * hl, Wb = q->w + 2 * hl;
  for (int m = 0; m < 4; ++m) {
    const int ph = s == 2 ? (m >> 1) : 0, pw = s == 2 ? (m & 1) : 0;
    uint64_t dims[4] = {static_cast<uint64_t>(fold == 1 ? CH : q->c), static_cast<uint64_t>(std::max(1, (Wb - pw + s - 1) / s)),
                        static_cast<uint64_t>(std::max(1, (Hb - ph + s - 1) / s)), static_cast<uint64_t>(q->n)};
    uint64_t strides[3] = {static_cast<uint64_t>(s) * q->c * es, static_cast<uint64_t>(s) * Wb * q->c * es,
                           static_cast<uint64_t>(Hb) * Wb * q->c * es};
    uint8_t* base = reinterpret_cast<uint8_t*>(q->ptr) + (static_cast<size_t>(ph) * Wb + pw) * q->c * es;
    rc = encode_tiled(&p.tmQ[m], q->dtype, 4, base, dims, strides, qbox, tf32 ? 2 : 1);
    if (rc != DTG_OK) return rc;
    if (s == 1) {
      for (int k = 1; k < 4; ++k) p.tmQ[k] = p.tmQ[0];
      break;
    }
  }

  static bool attr_set[2] = {false, false};
  static std::mutex mu;
  {
    std::lock_guard<std::mutex> lk(mu);
    if (!attr_set[tf32 ? 1 : 0]) {
      if (tf32)
        DTG_CHECK_CUDA(cudaFuncSetAttribute(wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      else
        DTG_CHECK_CUDA(cudaFuncSetAttribute(wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      attr_set[tf32 ? 1 : 0] = true;
    }
  }
  const size_t smem = static_cast<size_t>(pl.stages) * pl.stage_bytes + 1024 + 1024 + pl.kp * 128;
  dim3 grid(pl.splits, pl.ngroups, pl.mblocks);
  if (tf32)
    DTG_CHECK_CUDA(launch_k(wgrad_kernel<true>, grid, kWThreads, smem, stream, p));
  else
    DTG_CHECK_CUDA(launch_k(wgrad_kernel<false>, grid, kWThreads, smem, stream, p));
  const int rrows = p.rows_valid;
  const int ncv4 = (p.cols_valid + 3) / 4;
  const int total = pl.ntaps * rrows * ncv4;
  static const int rsub = getenv("DTG_REDUCE_SUB") ? atoi(getenv("DTG_REDUCE_SUB")) : 4;
#define DTG_REDUCE(SUB)                                                                                                        \
  DTG_CHECK_CUDA(launch_k(wgrad_reduce_kernel<SUB>, static_cast<int>((static_cast<long long>(SUB) * total + 255) / 256), 256, 0, \
                          stream, reinterpret_cast<const float4*>(p.ws), dw, pl.splits, pl.ntaps, p.mtot, pl.n_umma, a->pa, \
                          a->qb, fold, a->kw, 16 / es, rrows, ncv4, total, p.dbg))
  if (rsub == 1) {
    DTG_REDUCE(1);
  } else if (rsub == 4) {
    DTG_REDUCE(4);
  } else if (rsub == 8) {
    DTG_REDUCE(8);
  } else {
    DTG_REDUCE(16);
  }
#undef DTG_REDUCE
  return DTG_OK;
}
