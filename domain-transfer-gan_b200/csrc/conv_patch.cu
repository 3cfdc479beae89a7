// Patch-resident implicit-GEMM convolution (stride-1 forward and stride-1 dgrad) for sm_100a.
//
// The per-tap kernel (conv_igemm.cu) re-reads the activation tile from L2 once per filter tap and the
// weights once per tile; at ~42 B/clk/SM of L2->smem bandwidth that starves tcgen05 for every layer
// except the 128-channel residual convs.  Here an output tile is 8 x 16 pixels of one image and its
// whole input footprint -- the (8+kw-1) x (16+kh-1) pixel patch, one 128-byte channel chunk per pixel
// row -- is loaded by ONE TMA box per channel chunk (out-of-bounds = zero padding).  Every filter tap is
// then the same smem buffer seen through a UMMA descriptor whose start address is shifted by
// (dh * PW + dw) rows and whose 8-row-group stride (SBO) is the patch pitch PW * 128 B: SWIZZLE_128B is
// a function of the absolute smem address, so a shifted window reads exactly the rows TMA wrote
// (tools/umma_offset_test.cu).  The packed weights of all taps stay resident in smem for the lifetime of
// the persistent CTA (rows of 32/64/128 B with the matching swizzle), so the steady-state L2 traffic is
// the patch alone: 1.4x (3x3) .. 2.4x (7x7) the tile instead of taps x.
// Warp roles as in conv_igemm.cu: warp0 TMA producer, warp1 MMA issuer (+TMEM alloc), warps2-5 epilogue.
#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "conv_epilogue.cuh"

namespace dtg {

constexpr int kMaxMma = 200;
constexpr int kMaxAcc = 8;          // TMEM accumulator buffers (tile pipelining depth between MMA and epilogue)
constexpr int kPW = 8, kPH = 16;   // output tile (pixels); 8 wide so that one 8-row UMMA group = one tile row

struct PconvParams {
  CUtensorMap tmA, tmB;
  int PW, PH;            // patch extents (pixels)
  int a_stage_bytes;     // PW*PH*rb rounded up to 1024
  int org_h, org_w;      // patch origin = tile origin + org (plane-buffer coordinates)
  int tiles_w, tiles_h, N;
  int OHp, OWp, oh0, ow0;
  int ntaps;
  // per-MMA descriptor offsets (16-byte units) relative to the patch stage / the chunk's weight block, in issue
  // order (tap-major, k-step minor): read with uniform constant loads so that the single issuing thread spends
  // ~2 independent UIADDs per tcgen05.mma instead of a dependent address computation (tools/umma_rate_test.cu:
  // an M=128 N<=32 MMA occupies the pipe for only ~40 clk)
  unsigned short a_off[kMaxMma], b_off[kMaxMma];
  int nmma;              // MMAs per channel chunk = ntaps * ksteps
  int kchunks;
  int n_umma, rb, layout, b_tap_bytes, rb_elems;   // rb: bytes per smem row of BOTH operands (32 / 64 / 128)
  int a_stages, tmem_cols, nacc, nacc_log2;
  int dbg;   // DTG_PCONV_DBG experiments: 1 = skip MMAs, 2 = skip TMA patch loads, 4 = skip epilogue stores
  EpiParams e;
};

template <bool TF32>
__global__ void __launch_bounds__(kThreads, 1) pconv_kernel(const __grid_constant__ PconvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  pdl_trigger();      // the next kernel may start its prologue; it still waits for this grid before touching memory
  const int S = p.a_stages;
  uint8_t* sA = smem;
  uint8_t* sB = smem + S * p.a_stage_bytes;
  const int b_bytes = p.kchunks * p.ntaps * p.b_tap_bytes;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(sB + ((b_bytes + 1023) & ~1023));
  uint64_t* bar_empty = bar_full + S;
  uint64_t* bar_tfull = bar_empty + S;
  uint64_t* bar_tempty = bar_tfull + kMaxAcc;
  uint64_t* bar_b = bar_tempty + kMaxAcc;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_b + 1);
  uint8_t* epi_smem = reinterpret_cast<uint8_t*>(bar_full) + 1024;   // barriers occupy the first kBarrierBytes of this KB

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    if (p.e.use_tma) tma_prefetch_desc(&p.e.tmOut);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < S; ++i) {
        mbar_init(&bar_full[i], 1);
        mbar_init(&bar_empty[i], 1);
      }
      for (int i = 0; i < p.nacc; ++i) {
        mbar_init(&bar_tfull[i], 1);
        mbar_init(&bar_tempty[i], 4);
      }
      mbar_init(bar_b, 1);
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

  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int total_tiles = tiles_per_img * p.N;
  const uint32_t a_tx = static_cast<uint32_t>(p.PW) * p.PH * p.rb;

  if (warp == 0) {
    // ===================== TMA producer =====================
    {
      if (elect_one()) mbar_expect_tx(bar_b, static_cast<uint32_t>(b_bytes));
      __syncwarp();
      for (int kc = 0; kc < p.kchunks; ++kc)
        for (int t = 0; t < p.ntaps; ++t)
          if (elect_one())
            tma_load_2d(sB + (kc * p.ntaps + t) * p.b_tap_bytes, &p.tmB, bar_b, kc * p.rb_elems, t * p.n_umma);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n = tile / tiles_per_img;
        const int r = tile - n * tiles_per_img;
        const int th = r / p.tiles_w, tw = r - th * p.tiles_w;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(&bar_empty[stage], phase ^ 1);
          if (p.dbg & 2) {
            if (elect_one()) mbar_arrive(&bar_full[stage]);
          } else if (elect_one()) {
            mbar_expect_tx(&bar_full[stage], a_tx);
            tma_load_4d(sA + stage * p.a_stage_bytes, &p.tmA, &bar_full[stage], kc * p.rb_elems, tw * kPW + p.org_w,
                        th * kPH + p.org_h, n);
          }
          __syncwarp();
          if (++stage == S) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp walks the loop with warp-uniform operands (kernel parameters, uniform loop counters) and
    // only the tcgen05 instruction itself is predicated on elect.sync: descriptors then live in uniform
    // registers.  (Issuing from inside an `if (lane == 0)` region makes ptxas wrap every UTCHMMA in a
    // per-thread R2UR "waterfall" loop, ~100 clk per MMA -- more than an N <= 64 MMA occupies the tensor pipe.)
    const uint32_t idesc = umma_idesc(TF32 ? 2u : 1u, 0u, 0u, kTileM, p.n_umma);
    const uint32_t a_hi = ((static_cast<uint32_t>(p.PW) * p.rb) >> 4) | (1u << 14) | (static_cast<uint32_t>(p.layout) << 29);
    const uint32_t b_hi = ((8u * p.rb) >> 4) | (1u << 14) | (static_cast<uint32_t>(p.layout) << 29);
    const uint32_t b_tap16 = static_cast<uint32_t>(p.b_tap_bytes) >> 4;
    mbar_wait(bar_b, 0);
    tc_fence_after();
    const uint32_t b_lo0 = ((smem_u32(sB) >> 4) & 0x3FFFu) | (1u << 16);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int buf = it & (p.nacc - 1);
      const uint32_t use = static_cast<uint32_t>(it >> p.nacc_log2);
      mbar_wait(&bar_tempty[buf], (use & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * p.n_umma;
      for (int kc = 0; kc < p.kchunks; ++kc) {
        mbar_wait(&bar_full[stage], phase);
        tc_fence_after();
        const uint32_t a_lo0 = ((smem_u32(sA + stage * p.a_stage_bytes) >> 4) & 0x3FFFu) | (1u << 16);
        const uint32_t b_lok = b_lo0 + kc * p.ntaps * b_tap16;
        if (elect_one()) {
          // groups of 8: all 16 offsets are fetched (independent uniform constant loads) before the burst
          uint32_t acc = kc > 0 ? 1u : 0u;
          int i = 0;
          for (; i + 8 <= p.nmma; i += 8) {
            uint32_t ao[8], bo[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              ao[u] = a_lo0 + p.a_off[i + u];
              bo[u] = b_lok + p.b_off[i + u];
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              tc_mma<TF32>(d_tmem, (static_cast<uint64_t>(a_hi) << 32) | ao[u], (static_cast<uint64_t>(b_hi) << 32) | bo[u],
                           idesc, acc);
              acc = 1u;
            }
          }
          for (; i < p.nmma; ++i) {
            tc_mma<TF32>(d_tmem, (static_cast<uint64_t>(a_hi) << 32) | (a_lo0 + p.a_off[i]),
                         (static_cast<uint64_t>(b_hi) << 32) | (b_lok + p.b_off[i]), idesc, acc);
            acc = 1u;
          }
          tc_commit(&bar_empty[stage]);
        }
        __syncwarp();
        if (++stage == S) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) tc_commit(&bar_tfull[buf]);
      __syncwarp();
      ++it;
    }
  } else {
    // ===================== epilogue =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int iw = row % kPW, ih = row / kPW;
    uint8_t* stile = epi_smem + quad * kEpiWarpBytes;
    uint8_t* sbase = epi_smem + quad * kEpiTmaWarpBytes;     // (use_tma) the two layouts alias: only one is used per launch
    uint32_t cnt = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n = tile / tiles_per_img;
      const int r = tile - n * tiles_per_img;
      const int th = r / p.tiles_w, tw = r - th * p.tiles_w;
      const int a = th * kPH + ih, b = tw * kPW + iw;
      const bool valid = a < p.OHp && b < p.OWp;
      const int oh = p.oh0 + a, ow = p.ow0 + b;
      const int buf = it & (p.nacc - 1);
      const uint32_t use = static_cast<uint32_t>(it >> p.nacc_log2);
      if (!p.e.use_tma) epilogue_prepare<TF32>(p.e, valid, n, oh, ow, stile, lane);
      mbar_wait(&bar_tfull[buf], use & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + buf * p.n_umma;
      if (p.dbg & 4) {
      } else if (p.e.use_tma) {
        epilogue_tma<TF32>(p.e, taddr, tw * kPW, th * kPH + 4 * quad, n, sbase, cnt, lane, p.dbg);
      } else {
        epilogue_rows<TF32>(p.e, taddr, valid, n, oh, ow, stile, lane);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_tempty[buf]);
      ++it;
    }
    if (p.e.use_tma && cnt > 0) {
      if (elect_one()) bulk_wait_read<0>();     // staging smem must outlive the last stores' reads
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

static int pow2ceil_i(int v) {
  int r = 1;
  while (r < v) r <<= 1;
  return r;
}

int try_launch_pconv(const IgemmParams& g, const dtg_plane* in, const void* w, int w_rows, int w_cols, int taps_total,
                     int fold_w, cudaStream_t stream) {
  static const bool disabled = getenv("DTG_NO_PCONV") != nullptr;
  if (disabled || g.num_phases != 1 || g.out_step != 1) return 1;
  const int ntaps = g.ph_tap_begin[1];
  for (int t = 0; t < ntaps; ++t)
    if (g.tap_map[t] != 0) return 1;       // stride-2 parity maps: per-tap kernel
  const int es = elem_size(in->dtype);
  const bool tf32 = in->dtype == DTG_F32;
  const int kbytes = fold_w ? w_cols * es : std::min(w_cols, in->c) * es;
  const int kchunks = (kbytes + kRowBytes - 1) / kRowBytes;
  const int rb = kchunks > 1 ? 128 : std::max(32, pow2ceil_i(kbytes));
  const int b_tap_bytes = g.n_umma * rb;
  const int b_bytes = kchunks * ntaps * b_tap_bytes;
  int dh_min = 1 << 20, dh_max = -(1 << 20), dw_min = 1 << 20, dw_max = -(1 << 20);
  for (int t = 0; t < ntaps; ++t) {
    dh_min = std::min<int>(dh_min, g.tap_dh[t]);
    dh_max = std::max<int>(dh_max, g.tap_dh[t]);
    dw_min = std::min<int>(dw_min, g.tap_dw[t]);
    dw_max = std::max<int>(dw_max, g.tap_dw[t]);
  }
  const int PW = kPW + dw_max - dw_min, PH = kPH + dh_max - dh_min;
  if (PW > 256 || PH > 256) return 1;
  const int a_stage_bytes = (PW * PH * rb + 1023) & ~1023;
  const int fixed = 1024 + ((b_bytes + 1023) & ~1023) + 1024 + 4 * std::max(kEpiWarpBytes, kEpiTmaWarpBytes);
  const int budget = tensor_smem_budget() - fixed;
  if (b_bytes > 120 * 1024 || budget < 2 * a_stage_bytes) return 1;
  const int OHp = g.ph_OH[0], OWp = g.ph_OW[0];
  if (OWp < kPW || OHp < 4) return 1;      // tiny images: the batch-tiled per-tap kernel wastes less

  PconvParams p;
  memset(&p, 0, sizeof(p));
  p.PW = PW;
  p.PH = PH;
  p.a_stage_bytes = a_stage_bytes;
  p.org_h = dh_min;
  p.org_w = dw_min;
  p.tiles_w = (OWp + kPW - 1) / kPW;
  p.tiles_h = (OHp + kPH - 1) / kPH;
  p.N = g.N;
  p.OHp = OHp;
  p.OWp = OWp;
  p.oh0 = g.ph_oh0[0];
  p.ow0 = g.ph_ow0[0];
  p.ntaps = ntaps;
  const int ks = kchunks > 1 ? 4 : (kbytes + 31) / 32;
  if (kchunks > 1 && kbytes % kRowBytes != 0) return 1;
  if (ntaps * ks > kMaxMma) return 1;
  p.nmma = ntaps * ks;
  for (int t = 0; t < ntaps; ++t)
    for (int j = 0; j < ks; ++j) {
      const int row = (g.tap_dh[t] - dh_min) * PW + (g.tap_dw[t] - dw_min);
      p.a_off[t * ks + j] = static_cast<unsigned short>((row * rb + j * 32) >> 4);
      p.b_off[t * ks + j] = static_cast<unsigned short>((g.tap_w[t] * b_tap_bytes + j * 32) >> 4);
    }
  p.kchunks = kchunks;
  p.n_umma = g.n_umma;
  p.rb = rb;
  p.layout = rb == 128 ? 2 : (rb == 64 ? 4 : 6);
  p.b_tap_bytes = b_tap_bytes;
  p.rb_elems = rb / es;
  p.a_stages = std::max(2, std::min(4, budget / a_stage_bytes));
  p.nacc_log2 = 1;
  while ((2 << p.nacc_log2) <= kMaxAcc && (2 << p.nacc_log2) * g.n_umma <= 512) ++p.nacc_log2;
  p.nacc = 1 << p.nacc_log2;
  p.tmem_cols = 32;
  while (p.tmem_cols < p.nacc * g.n_umma) p.tmem_cols <<= 1;
  p.e = g.e;
  p.e.use_tma = 0;
  {
    const int rowb = std::min(p.e.out_C * es, 128);
    const bool pow2 = rowb == 32 || rowb == 64 || rowb == 128;
    static const bool no_tma_epi = getenv("DTG_NO_TMA_EPI") != nullptr;
    if (!no_tma_epi && !p.e.out_nchw && !p.e.out_reflect && p.e.act != DTG_ACT_TANH && pow2 && (p.e.out_C * es) % rowb == 0) {
      const int oh_ = p.e.out_halo, Hb_ = p.e.out_H + 2 * oh_, Wb_ = p.e.out_W + 2 * oh_;
      uint8_t* base = reinterpret_cast<uint8_t*>(p.e.out) +
                      (static_cast<size_t>(p.oh0 + oh_) * Wb_ + (p.ow0 + oh_)) * p.e.out_C * es;
      uint64_t dims[4] = {static_cast<uint64_t>(p.e.out_C), static_cast<uint64_t>(OWp), static_cast<uint64_t>(OHp),
                          static_cast<uint64_t>(g.N)};
      uint64_t strides[3] = {static_cast<uint64_t>(p.e.out_C) * es, static_cast<uint64_t>(Wb_) * p.e.out_C * es,
                             static_cast<uint64_t>(Hb_) * Wb_ * p.e.out_C * es};
      uint32_t box[4] = {static_cast<uint32_t>(rowb / es), static_cast<uint32_t>(kPW), 4u, 1u};
      int rc = encode_tiled(&p.e.tmOut, in->dtype, 4, base, dims, strides, box, rowb == 128 ? 1 : (rowb == 64 ? 3 : 4));
      if (rc != DTG_OK) return rc;
      p.e.use_tma = 1;
      p.e.row_bytes = rowb;
    }
  }
  {
    const char* d = getenv("DTG_PCONV_DBG");
    p.dbg = d ? atoi(d) : 0;
    if (p.dbg & 1) p.nmma = 1;
    if (p.dbg & 8) p.a_stages = 2;
  }

  const int hl = in->halo;
  const int Hb = in->h + 2 * hl, Wb = in->w + 2 * hl;
  {
    // fold_w: overlapping rows -- pixel p's row is the 128 bytes of pixels p..p+7 (dim-1 stride = 16 B < row size)
    uint64_t dims[4] = {static_cast<uint64_t>(fold_w ? rb / es : in->c), static_cast<uint64_t>(Wb), static_cast<uint64_t>(Hb),
                        static_cast<uint64_t>(in->n)};
    uint64_t strides[3] = {static_cast<uint64_t>(in->c) * es, static_cast<uint64_t>(Wb) * in->c * es,
                           static_cast<uint64_t>(Hb) * Wb * in->c * es};
    uint32_t box[4] = {static_cast<uint32_t>(rb / es), static_cast<uint32_t>(PW), static_cast<uint32_t>(PH), 1u};
    int rc = encode_tiled(&p.tmA, in->dtype, 4, in->ptr, dims, strides, box, rb == 128 ? 1 : (rb == 64 ? 3 : 4));
    if (rc != DTG_OK) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(w_cols), static_cast<uint64_t>(w_rows) * taps_total};
    uint64_t strides[1] = {static_cast<uint64_t>(w_cols) * es};
    uint32_t box[2] = {static_cast<uint32_t>(rb / es), static_cast<uint32_t>(w_rows)};
    int rc = encode_tiled(&p.tmB, in->dtype, 2, const_cast<void*>(w), dims, strides, box, rb == 128 ? 1 : (rb == 64 ? 3 : 4));
    if (rc != DTG_OK) return rc;
  }

  static int num_sms = 0;
  static bool attr_set[2] = {false, false};
  static std::mutex mu;
  {
    std::lock_guard<std::mutex> lk(mu);
    if (num_sms == 0) {
      int dev = 0;
      DTG_CHECK_CUDA(cudaGetDevice(&dev));
      DTG_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    if (!attr_set[tf32 ? 1 : 0]) {
      if (tf32)
        DTG_CHECK_CUDA(cudaFuncSetAttribute(pconv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      else
        DTG_CHECK_CUDA(cudaFuncSetAttribute(pconv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      attr_set[tf32 ? 1 : 0] = true;
    }
  }
  const size_t smem = static_cast<size_t>(fixed) + static_cast<size_t>(p.a_stages) * a_stage_bytes;
  const int total = p.tiles_w * p.tiles_h * p.N;
  const int grid = std::max(1, std::min(total, num_sms));
  if (tf32)
    DTG_CHECK_CUDA(launch_k(pconv_kernel<true>, grid, kThreads, smem, stream, p));
  else
    DTG_CHECK_CUDA(launch_k(pconv_kernel<false>, grid, kThreads, smem, stream, p));
  return DTG_OK;
}

}  // namespace dtg
