// Patch-resident implicit-GEMM convolution, streamed-weight variant for the 128-channel residual stack
// (modules.py:162,180,211,227 and their data gradients): 3x3, stride 1, 128 -> 128 channels.
//
// Its packed weights (9 taps x 128 x 128 bf16 = 295 KB) do not fit in shared memory, and re-streaming them from L2 for
// every 128-pixel tile (as the per-tap kernel does, together with one activation tile per tap: 590 KB per tile)
// pins that kernel at the ~50 B/clk/SM L2->SM ceiling.  Here a work item is TWO 8x16-pixel tiles: their input
// patches (one TMA box per tile and 64-channel chunk, every tap a row-shifted UMMA window, as in conv_patch.cu) stay
// resident while each (chunk, tap) weight block streams through a small ring ONCE and feeds both tiles' accumulators:
// (2 x 46 KB + 295 KB) / 2 = 194 KB of L2 traffic per tile, under the ceiling for 72 MMAs of 64 clk.
// Flat-raster mode (the data gradient w.r.t. a reflect-padded input, ring > 0): dy carries a ZERO halo ring equal to the
// ring, so its buffer [n][h+2r][w+2r][c] has the same raster as the output buffer.  A tile is then 128 CONSECUTIVE buffer
// pixels of the tall n*(h+2r) x (w+2r) image (722 tiles instead of 1200 for 80 x 34 x 34), its patch the 128 + 2*(w+2r+1)
// pixels around it (one 2-D TMA box, out-of-range = zero fill), every tap the window shifted by dh*(w+2r) + dw pixels
// (row wrap-around lands in the zero ring = the padding), 8-row groups contiguous, and the tile leaves as 4 x 32 pixels.
// Warp roles: warp0 TMA producer, warp1 MMA issuer (+TMEM alloc), warps2-5 epilogue; two TMEM accumulator sets.
#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "conv_epilogue.cuh"

namespace dtg {

constexpr int kP2W = 8, kP2H = 16;
constexpr int kP2MaxUnits = 8, kP2MaxB = 8;

struct Pconv2Params {
  CUtensorMap tmA, tmB;
  int PW, PH, a_unit_bytes;
  int org_h, org_w;
  int tiles_w, tiles_h, total_tiles;
  int pairs, items;      // items [0, pairs) are two tiles, the rest single tiles (tail balancing)
  int OHp, OWp, oh0, ow0;
  int ntaps;
  unsigned short a_off[kMaxTaps];   // tap window start inside a patch unit, 16-byte units
  unsigned char tap_w[kMaxTaps];
  int kchunks, n_umma, b_stage_bytes;
  int a_units, b_stages, tmem_cols;
  int flat;              // flat-raster mode
  int lead;              // flat: patch starts `lead` pixels before the tile's first pixel
  int a_tx;              // bytes of one patch unit
  int a_sbo;             // byte stride between 8-row groups of the A window (patch pitch * 128, flat: 1024)
  int dbg;   // DTG_P2_DBG experiments: 1 skip weight loads, 2 skip patch loads, 4 skip epilogue
  EpiParams e;
};

template <bool TF32>
__global__ void __launch_bounds__(kThreads, 1) pconv2_kernel(const __grid_constant__ Pconv2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  pdl_trigger();
  constexpr int KC = TF32 ? 32 : 64;
  const int AU = p.a_units, BS = p.b_stages;
  uint8_t* sA = smem;
  uint8_t* sB = smem + AU * p.a_unit_bytes;
  uint64_t* bar_afull = reinterpret_cast<uint64_t*>(sB + BS * p.b_stage_bytes);
  uint64_t* bar_aempty = bar_afull + kP2MaxUnits;
  uint64_t* bar_bfull = bar_aempty + kP2MaxUnits;
  uint64_t* bar_bempty = bar_bfull + kP2MaxB;
  uint64_t* bar_tfull = bar_bempty + kP2MaxB;
  uint64_t* bar_tempty = bar_tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tempty + 2);
  uint8_t* epi_smem = reinterpret_cast<uint8_t*>(bar_afull) + 1024;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    if (p.e.use_tma) tma_prefetch_desc(&p.e.tmOut);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < AU; ++i) {
        mbar_init(&bar_afull[i], 1);
        mbar_init(&bar_aempty[i], 1);
      }
      for (int i = 0; i < BS; ++i) {
        mbar_init(&bar_bfull[i], 1);
        mbar_init(&bar_bempty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bar_tfull[i], 1);
        mbar_init(&bar_tempty[i], 4);
      }
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, p.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int items = p.items;
  const uint32_t a_tx = static_cast<uint32_t>(p.a_tx);
  const uint32_t b_tx = static_cast<uint32_t>(p.b_stage_bytes);

  if (warp == 0) {
    // ===================== TMA producer =====================
    int au = 0, bs = 0;
    uint32_t aph = 0, bph = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int nt = item < p.pairs ? 2 : 1;
      const int tile0 = item < p.pairs ? 2 * item : p.pairs + item;
      for (int kc = 0; kc < p.kchunks; ++kc) {
        for (int mt = 0; mt < nt; ++mt) {
          const int tile = tile0 + mt;
          const int n = tile / tiles_per_img;
          const int r = tile - n * tiles_per_img;
          const int th = r / p.tiles_w, tw = r - th * p.tiles_w;
          mbar_wait(&bar_aempty[au], aph ^ 1);
          if (p.dbg & 2) {
            if (elect_one()) mbar_arrive(&bar_afull[au]);
          } else if (elect_one()) {
            mbar_expect_tx(&bar_afull[au], a_tx);
            if (p.flat)
              tma_load_2d(sA + au * p.a_unit_bytes, &p.tmA, &bar_afull[au], kc * KC, tile * kTileM - p.lead);
            else
              tma_load_4d(sA + au * p.a_unit_bytes, &p.tmA, &bar_afull[au], kc * KC, tw * kP2W + p.org_w, th * kP2H + p.org_h, n);
          }
          __syncwarp();
          if (++au == AU) {
            au = 0;
            aph ^= 1;
          }
        }
        for (int t = 0; t < p.ntaps; ++t) {
          mbar_wait(&bar_bempty[bs], bph ^ 1);
          if (p.dbg & 1) {
            if (elect_one()) mbar_arrive(&bar_bfull[bs]);
          } else if (elect_one()) {
            mbar_expect_tx(&bar_bfull[bs], b_tx);
            tma_load_2d(sB + bs * p.b_stage_bytes, &p.tmB, &bar_bfull[bs], kc * KC, p.tap_w[t] * p.n_umma);
          }
          __syncwarp();
          if (++bs == BS) {
            bs = 0;
            bph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform; tcgen05 instructions behind elect.sync) =====================
    const uint32_t idesc = umma_idesc(TF32 ? 2u : 1u, 0u, 0u, kTileM, p.n_umma);
    const uint32_t a_hi = (static_cast<uint32_t>(p.a_sbo) >> 4) | (1u << 14) | (2u << 29);
    const uint32_t b_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    int au = 0, bs = 0;
    uint32_t aph = 0, bph = 0;
    int it = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int nt = item < p.pairs ? 2 : 1;
      const int set = it & 1;
      mbar_wait(&bar_tempty[set], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d0 = tmem_base + set * 2 * p.n_umma;
      for (int kc = 0; kc < p.kchunks; ++kc) {
        uint32_t a_lo[2];
        int unit[2];
        for (int mt = 0; mt < nt; ++mt) {
          mbar_wait(&bar_afull[au], aph);
          unit[mt] = au;
          a_lo[mt] = ((smem_u32(sA + au * p.a_unit_bytes) >> 4) & 0x3FFFu) | (1u << 16);
          if (++au == AU) {
            au = 0;
            aph ^= 1;
          }
        }
        if (nt == 1) a_lo[1] = a_lo[0];
        tc_fence_after();
        for (int t = 0; t < p.ntaps; ++t) {
          mbar_wait(&bar_bfull[bs], bph);
          tc_fence_after();
          const uint32_t b_lo = ((smem_u32(sB + bs * p.b_stage_bytes) >> 4) & 0x3FFFu) | (1u << 16);
          const uint32_t ao = p.a_off[t];
          const uint32_t acc0 = (kc > 0 || t > 0) ? 1u : 0u;
          if (elect_one()) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              tc_mma<TF32>(d0, (static_cast<uint64_t>(a_hi) << 32) | (a_lo[0] + ao + 2 * j),
                           (static_cast<uint64_t>(b_hi) << 32) | (b_lo + 2 * j), idesc, j > 0 ? 1u : acc0);
            if (nt == 2) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                tc_mma<TF32>(d0 + p.n_umma, (static_cast<uint64_t>(a_hi) << 32) | (a_lo[1] + ao + 2 * j),
                             (static_cast<uint64_t>(b_hi) << 32) | (b_lo + 2 * j), idesc, j > 0 ? 1u : acc0);
            }
            tc_commit(&bar_bempty[bs]);
          }
          __syncwarp();
          if (++bs == BS) {
            bs = 0;
            bph ^= 1;
          }
        }
        if (elect_one()) {
          tc_commit(&bar_aempty[unit[0]]);
          if (nt == 2) tc_commit(&bar_aempty[unit[1]]);
        }
        __syncwarp();
      }
      if (elect_one()) tc_commit(&bar_tfull[set]);
      __syncwarp();
      ++it;
    }
  } else {
    // ===================== epilogue =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int iw = row % kP2W, ih = row / kP2W;
    uint8_t* stile = epi_smem + quad * kEpiWarpBytes;
    uint8_t* sbase = epi_smem + quad * kEpiTmaWarpBytes;
    uint32_t cnt = 0;
    int it = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int nt = item < p.pairs ? 2 : 1;
      const int tile0 = item < p.pairs ? 2 * item : p.pairs + item;
      const int set = it & 1;
      mbar_wait(&bar_tfull[set], (it >> 1) & 1);
      tc_fence_after();
      for (int mt = 0; mt < ((p.dbg & 4) ? 0 : nt); ++mt) {
        const int tile = tile0 + mt;
        const int n = tile / tiles_per_img;
        const int r = tile - n * tiles_per_img;
        const int th = r / p.tiles_w, tw = r - th * p.tiles_w;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + (set * 2 + mt) * p.n_umma;
        if (p.flat) {      // tmOut views the output buffer as [c][8][pixels / 8]: this warp's 32 pixels are 4 rows of it
          epilogue_tma<TF32>(p.e, taddr, 0, (tile * kTileM + 32 * quad) >> 3, 0, sbase, cnt, lane);
        } else if (p.e.use_tma) {
          epilogue_tma<TF32>(p.e, taddr, tw * kP2W, th * kP2H + 4 * quad, n, sbase, cnt, lane);
        } else {
          const int a = th * kP2H + ih, b = tw * kP2W + iw;
          const bool valid = a < p.OHp && b < p.OWp;
          const int oh = p.oh0 + a, ow = p.ow0 + b;
          epilogue_prepare<TF32>(p.e, valid, n, oh, ow, stile, lane);
          epilogue_rows<TF32>(p.e, taddr, valid, n, oh, ow, stile, lane);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_tempty[set]);
      ++it;
    }
    if (p.e.use_tma && cnt > 0) {
      if (elect_one()) bulk_wait_read<0>();
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

int try_launch_pconv2(const IgemmParams& g, const dtg_plane* in, const void* w, int w_rows, int w_cols, int taps_total,
                      cudaStream_t stream) {
  static const bool disabled = getenv("DTG_NO_PCONV2") != nullptr || getenv("DTG_NO_PCONV") != nullptr;
  if (disabled || g.num_phases != 1 || g.out_step != 1) return 1;
  const int ntaps = g.ph_tap_begin[1];
  for (int t = 0; t < ntaps; ++t)
    if (g.tap_map[t] != 0) return 1;
  const int es = elem_size(in->dtype);
  const bool tf32 = in->dtype == DTG_F32;
  const int kbytes = std::min(w_cols, in->c) * es;
  if (kbytes % kRowBytes != 0 || g.n_umma > 128) return 1;      // full 128-byte chunks, two accumulator pairs in TMEM
  const int kchunks = kbytes / kRowBytes;
  int dh_min = 1 << 20, dh_max = -(1 << 20), dw_min = 1 << 20, dw_max = -(1 << 20);
  for (int t = 0; t < ntaps; ++t) {
    dh_min = std::min<int>(dh_min, g.tap_dh[t]);
    dh_max = std::max<int>(dh_max, g.tap_dh[t]);
    dw_min = std::min<int>(dw_min, g.tap_dw[t]);
    dw_max = std::max<int>(dw_max, g.tap_dw[t]);
  }
  const int OHp = g.ph_OH[0], OWp = g.ph_OW[0];
  // flat-raster mode: dy with a zero halo == ring, output plane with the same halo (identical buffer rasters)
  const int hd = g.flat_dgrad ? in->halo : 0;
  const bool flat = g.flat_dgrad != 0;
  const int Wr = in->w + 2 * hd, Hr = in->h + 2 * hd;
  const long long flat_px = static_cast<long long>(in->n) * Hr * Wr;
  int lead = 0;
  if (flat) {
    if (g.e.out_nchw || g.e.out_reflect || g.e.act == DTG_ACT_TANH || g.e.out_halo != hd || g.e.out_H != in->h || g.e.out_W != in->w ||
        g.e.out_C * es % kRowBytes != 0 || OHp != Hr || OWp != Wr || g.ph_oh0[0] != -hd || g.ph_ow0[0] != -hd || flat_px % 8 != 0 ||
        flat_px > (1ll << 30))
      return 1;
    for (int t = 0; t < ntaps; ++t) lead = std::max(lead, std::abs((g.tap_dh[t] + hd) * Wr + g.tap_dw[t] + hd));
    if (kTileM + 2 * lead > 256) return 1;
  }
  const int PW = flat ? 1 : kP2W + dw_max - dw_min, PH = flat ? kTileM + 2 * lead : kP2H + dh_max - dh_min;
  const int a_unit_bytes = (PW * PH * kRowBytes + 1023) & ~1023;
  const int b_stage_bytes = g.n_umma * kRowBytes;
  if (!flat) {
    if (OWp < kP2W || OHp < 8 || PW > 256 || PH > 256) return 1;
    if (OWp % kP2W != 0 || OHp % kP2H != 0) return 1;     // ragged 8x16 tilings (e.g. 34x34 ring outputs) waste > 40 % of the MMAs
  }
  const int epi = 4 * std::max(kEpiWarpBytes, kEpiTmaWarpBytes);
  const int budget = tensor_smem_budget() - 1024 - 1024 - epi;
  // ring sizes: at least the 2*kchunks patch units of one item plus one to prefetch; 3-4 weight stages
  // the weight ring must cover the L2 round trip (a stage is consumed in 8 MMAs = 512 clk): as many stages as fit
  // next to the 2*kchunks patch units of one item
  int a_units = 2 * kchunks;
  if (const char* e = getenv("DTG_P2_AUNITS")) a_units = atoi(e);
  int b_stages = std::min(kP2MaxB, (budget - a_units * a_unit_bytes) / b_stage_bytes);
  if (a_units < 2 * kchunks || a_units > kP2MaxUnits || b_stages < 3) return 1;

  Pconv2Params p;
  memset(&p, 0, sizeof(p));
  p.PW = PW;
  p.PH = PH;
  p.a_unit_bytes = a_unit_bytes;
  p.org_h = dh_min;
  p.org_w = dw_min;
  p.tiles_w = (OWp + kP2W - 1) / kP2W;
  p.tiles_h = (OHp + kP2H - 1) / kP2H;
  p.total_tiles = flat ? static_cast<int>((flat_px + kTileM - 1) / kTileM) : p.tiles_w * p.tiles_h * g.N;
  p.flat = flat ? 1 : 0;
  p.lead = lead;
  p.a_tx = PW * PH * kRowBytes;
  p.a_sbo = flat ? 1024 : PW * kRowBytes;
  p.OHp = OHp;
  p.OWp = OWp;
  p.oh0 = g.ph_oh0[0];
  p.ow0 = g.ph_ow0[0];
  p.ntaps = ntaps;
  for (int t = 0; t < ntaps; ++t) {
    const int rows = flat ? lead + (g.tap_dh[t] + hd) * Wr + g.tap_dw[t] + hd : (g.tap_dh[t] - dh_min) * PW + (g.tap_dw[t] - dw_min);
    p.a_off[t] = static_cast<unsigned short>((rows * kRowBytes) >> 4);
    p.tap_w[t] = g.tap_w[t];
  }
  p.kchunks = kchunks;
  p.n_umma = g.n_umma;
  p.b_stage_bytes = b_stage_bytes;
  p.a_units = a_units;
  p.b_stages = b_stages;
  p.tmem_cols = 32;
  while (p.tmem_cols < 4 * g.n_umma) p.tmem_cols <<= 1;
  p.e = g.e;
  p.e.use_tma = 0;
  if (const char* d = getenv("DTG_P2_DBG")) p.dbg = atoi(d);
  {
    const int rowb = std::min(p.e.out_C * es, 128);
    const bool pow2 = rowb == 32 || rowb == 64 || rowb == 128;
    static const bool no_tma_epi = getenv("DTG_NO_TMA_EPI") != nullptr;
    if (flat) {
      // the whole output buffer as [c][8][pixels / 8]: a {row, 8, 4, 1} box is 32 consecutive pixels
      uint64_t dims[4] = {static_cast<uint64_t>(p.e.out_C), 8ull, static_cast<uint64_t>(flat_px / 8), 1ull};
      uint64_t strides[3] = {static_cast<uint64_t>(p.e.out_C) * es, static_cast<uint64_t>(p.e.out_C) * es * 8,
                             static_cast<uint64_t>(p.e.out_C) * es * static_cast<uint64_t>(flat_px)};
      uint32_t box[4] = {static_cast<uint32_t>(rowb / es), 8u, 4u, 1u};
      int rc = encode_tiled(&p.e.tmOut, in->dtype, 4, p.e.out, dims, strides, box, 1);
      if (rc != DTG_OK) return rc;
      p.e.use_tma = 1;
      p.e.row_bytes = rowb;
    } else if (!no_tma_epi && !p.e.out_nchw && !p.e.out_reflect && p.e.act != DTG_ACT_TANH && pow2 && (p.e.out_C * es) % rowb == 0) {
      const int oh_ = p.e.out_halo, Hb_ = p.e.out_H + 2 * oh_, Wb_ = p.e.out_W + 2 * oh_;
      uint8_t* base = reinterpret_cast<uint8_t*>(p.e.out) + (static_cast<size_t>(p.oh0 + oh_) * Wb_ + (p.ow0 + oh_)) * p.e.out_C * es;
      uint64_t dims[4] = {static_cast<uint64_t>(p.e.out_C), static_cast<uint64_t>(OWp), static_cast<uint64_t>(OHp),
                          static_cast<uint64_t>(g.N)};
      uint64_t strides[3] = {static_cast<uint64_t>(p.e.out_C) * es, static_cast<uint64_t>(Wb_) * p.e.out_C * es,
                             static_cast<uint64_t>(Hb_) * Wb_ * p.e.out_C * es};
      uint32_t box[4] = {static_cast<uint32_t>(rowb / es), static_cast<uint32_t>(kP2W), 4u, 1u};
      int rc = encode_tiled(&p.e.tmOut, in->dtype, 4, base, dims, strides, box, rowb == 128 ? 1 : (rowb == 64 ? 3 : 4));
      if (rc != DTG_OK) return rc;
      p.e.use_tma = 1;
      p.e.row_bytes = rowb;
    }
  }

  const int hl = in->halo;
  const int Hb = in->h + 2 * hl, Wb = in->w + 2 * hl;
  if (flat) {
    uint64_t dims[2] = {static_cast<uint64_t>(in->c), static_cast<uint64_t>(flat_px)};
    uint64_t strides[1] = {static_cast<uint64_t>(in->c) * es};
    uint32_t box[2] = {static_cast<uint32_t>(kRowBytes / es), static_cast<uint32_t>(PH)};
    int rc = encode_tiled(&p.tmA, in->dtype, 2, in->ptr, dims, strides, box, 1);
    if (rc != DTG_OK) return rc;
  } else {
    uint64_t dims[4] = {static_cast<uint64_t>(in->c), static_cast<uint64_t>(Wb), static_cast<uint64_t>(Hb), static_cast<uint64_t>(in->n)};
    uint64_t strides[3] = {static_cast<uint64_t>(in->c) * es, static_cast<uint64_t>(Wb) * in->c * es,
                           static_cast<uint64_t>(Hb) * Wb * in->c * es};
    uint32_t box[4] = {static_cast<uint32_t>(kRowBytes / es), static_cast<uint32_t>(PW), static_cast<uint32_t>(PH), 1u};
    int rc = encode_tiled(&p.tmA, in->dtype, 4, in->ptr, dims, strides, box, 1);
    if (rc != DTG_OK) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(w_cols), static_cast<uint64_t>(w_rows) * taps_total};
    uint64_t strides[1] = {static_cast<uint64_t>(w_cols) * es};
    uint32_t box[2] = {static_cast<uint32_t>(kRowBytes / es), static_cast<uint32_t>(w_rows)};
    int rc = encode_tiled(&p.tmB, in->dtype, 2, const_cast<void*>(w), dims, strides, box, 1);
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
        DTG_CHECK_CUDA(cudaFuncSetAttribute(pconv2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      else
        DTG_CHECK_CUDA(cudaFuncSetAttribute(pconv2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      attr_set[tf32 ? 1 : 0] = true;
    }
  }
  const size_t smem = 1024 + static_cast<size_t>(a_units) * a_unit_bytes + static_cast<size_t>(b_stages) * b_stage_bytes + 1024 + epi;
  // complete rounds of two-tile items, then single tiles: the last round costs one tile instead of two
  const int grid = std::max(1, std::min((p.total_tiles + 1) / 2, num_sms));
  p.pairs = (p.total_tiles / (2 * grid)) * grid;
  p.items = p.pairs + (p.total_tiles - 2 * p.pairs);
  if (tf32)
    DTG_CHECK_CUDA(launch_k(pconv2_kernel<true>, grid, kThreads, smem, stream, p));
  else
    DTG_CHECK_CUDA(launch_k(pconv2_kernel<false>, grid, kThreads, smem, stream, p));
  return DTG_OK;
}

}  // namespace dtg
