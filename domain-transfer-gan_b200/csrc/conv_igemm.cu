// Implicit-GEMM convolution for sm_100a: TMA-staged NHWC tiles -> tcgen05.mma (TMEM accumulators).
//
// GEMM view:  D[pixel, cout] = sum_{tap, cin} A_tap[pixel, cin] * W[tap][cout][cin]
//   M tile  = up to 128 output pixels forming a rectangular patch (bw x bh x bn) so that, for every
//             filter tap, the A operand is ONE rectangular TMA box of the NHWC activation tensor
//             (out-of-bounds coordinates are zero-filled by TMA = zero padding for free; reflection
//             padding is materialised in the plane's halo by the producing kernel).
//   N       = all output channels (16..256, one UMMA instruction wide, accumulator in TMEM)
//   K loop  = taps x 128-byte channel chunks (64 bf16 / 32 tf32), 4 UMMA K-steps per chunk.
// Stride-2 forward convs read one of four parity sub-sampled tensor maps per tap; dgrad and
// ConvTranspose2d run as output-parity phases, each a dense stride-1 tap list (no zero insertion).
// Warp roles: warp0 = TMA producer, warp1 = MMA issuer (+TMEM alloc), warps2-5 = epilogue
// (TMEM -> registers -> bias/activation -> global).  Persistent CTAs, double-buffered accumulators.
#include <algorithm>
#include <mutex>
#include <type_traits>

#include "common.cuh"
#include "conv_epilogue.cuh"

namespace dtg {

template <bool TF32>
__global__ void __launch_bounds__(kThreads, 1) igemm_kernel(const __grid_constant__ IgemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  pdl_trigger();      // the next kernel may start its prologue; it still waits for this grid before touching memory
  constexpr int KC = TF32 ? 32 : 64;  // elements per 128-byte row
  const int b_bytes = p.n_umma * kRowBytes;
  const int stage_bytes = kATileBytes + b_bytes;
  const int S = p.stages;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + S * stage_bytes);
  uint64_t* bar_empty = bar_full + S;
  uint64_t* bar_tfull = bar_empty + S;
  uint64_t* bar_tempty = bar_tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tempty + 2);
  uint8_t* epi_smem = smem + S * stage_bytes + kBarrierBytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.tmA[i]);
    tma_prefetch_desc(&p.tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < S; ++i) {
        mbar_init(&bar_full[i], 1);
        mbar_init(&bar_empty[i], 1);
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
  pdl_wait();         // barriers / TMEM / descriptor prefetch above overlap the previous kernel's tail

  const int tiles_per_phase = p.tiles_w * p.tiles_h * p.tiles_n;
  const int total_tiles = tiles_per_phase * p.num_phases;
  const int box_rows = p.bw * p.bh * p.bn;
  const uint32_t tx_bytes = box_rows * kRowBytes + b_bytes;

  if (warp == 0) {
    // ===================== TMA producer =====================
    {   // warp-uniform loop; only the TMA / mbarrier instructions are predicated on elect.sync
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int ph = tile / tiles_per_phase;
        int r = tile - ph * tiles_per_phase;
        const int tw = r % p.tiles_w;
        r /= p.tiles_w;
        const int th = r % p.tiles_h;
        const int tn = r / p.tiles_h;
        const int a0 = th * p.bh, b0 = tw * p.bw, n0 = tn * p.bn;
        if (a0 >= p.ph_OH[ph] || b0 >= p.ph_OW[ph]) continue;
        for (int t = p.ph_tap_begin[ph]; t < p.ph_tap_begin[ph + 1]; ++t) {
          const CUtensorMap* mapA = &p.tmA[p.tap_map[t]];
          const int cw = b0 + p.tap_dw[t], chh = a0 + p.tap_dh[t];
          const int wrow = p.tap_w[t] * p.n_umma;
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait(&bar_empty[stage], phase ^ 1);
            uint8_t* sa = smem + stage * stage_bytes;
            if (elect_one()) {
              mbar_expect_tx(&bar_full[stage], tx_bytes);
              tma_load_4d(sa, mapA, &bar_full[stage], kc * KC, cw, chh, n0);
              tma_load_2d(sa + kATileBytes, &p.tmB, &bar_full[stage], kc * KC, wrow);
            }
            __syncwarp();
            if (++stage == S) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = umma_idesc(TF32 ? 2u : 1u, 0u, 0u, kTileM, p.n_umma);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int ph = tile / tiles_per_phase;
      int r = tile - ph * tiles_per_phase;
      const int tw = r % p.tiles_w;
      r /= p.tiles_w;
      const int th = r % p.tiles_h;
      if (th * p.bh >= p.ph_OH[ph] || tw * p.bw >= p.ph_OW[ph]) continue;
      const int buf = it & 1;
      const uint32_t use = static_cast<uint32_t>(it >> 1);
      mbar_wait(&bar_tempty[buf], (use & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * p.n_umma;
      const int nk = (p.ph_tap_begin[ph + 1] - p.ph_tap_begin[ph]) * p.kchunks;
      for (int k = 0; k < nk; ++k) {
        mbar_wait(&bar_full[stage], phase);
        tc_fence_after();
        {
          const uint32_t sa = smem_u32(smem + stage * stage_bytes);
          const uint32_t sb = sa + kATileBytes;
          const uint64_t ad0 = umma_desc_sw128(sa, 16, 1024);
          const uint64_t bd0 = umma_desc_sw128(sb, 16, 1024);
          if (elect_one()) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (j < p.ksteps) tc_mma<TF32>(d_tmem, ad0 + 2 * j, bd0 + 2 * j, idesc, (k > 0 || j > 0) ? 1u : 0u);
            tc_commit(&bar_empty[stage]);
          }
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
    // TMEM -> registers -> (bias, activation, convert) -> per-warp smem staging (128-byte channel chunks) ->
    // global rows written as full 128-byte lines (8 lanes x 16 B per row, 4 rows per warp instruction).
    const int quad = warp & 3;           // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;
    const int iw = row % p.bw;
    const int ih = (row / p.bw) % p.bh;
    const int in_ = row / (p.bw * p.bh);
    uint8_t* stile = epi_smem + quad * kEpiWarpBytes;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int ph = tile / tiles_per_phase;
      int r = tile - ph * tiles_per_phase;
      const int tw = r % p.tiles_w;
      r /= p.tiles_w;
      const int th = r % p.tiles_h;
      const int tn = r / p.tiles_h;
      const int a0 = th * p.bh, b0 = tw * p.bw, n0 = tn * p.bn;
      if (a0 >= p.ph_OH[ph] || b0 >= p.ph_OW[ph]) continue;
      const int buf = it & 1;
      const uint32_t use = static_cast<uint32_t>(it >> 1);
      const int a = a0 + ih, b = b0 + iw, n = n0 + in_;
      const bool valid = (row < box_rows) && a < p.ph_OH[ph] && b < p.ph_OW[ph] && n < p.N;
      const int oh = p.ph_oh0[ph] + a * p.out_step;
      const int ow = p.ph_ow0[ph] + b * p.out_step;
      epilogue_prepare<TF32>(p.e, valid, n, oh, ow, stile, lane);
      mbar_wait(&bar_tfull[buf], use & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + buf * p.n_umma;
      epilogue_rows<TF32>(p.e, taddr, valid, n, oh, ow, stile, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_tempty[buf]);
      ++it;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// --------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------
static void choose_patch(int PW, int PH, int N, int* bw, int* bh, int* bn) {
  // maximise useful pixels per 128-row MMA tile; prefer wider boxes on ties
  double best = -1.0;
  int bbw = 1, bbh = 1;
  for (int w = std::min(PW, 128); w >= 1; --w) {
    int hmax = std::min(PH, 128 / w);
    for (int h = hmax; h >= 1; --h) {
      long tiles = static_cast<long>((PW + w - 1) / w) * ((PH + h - 1) / h);
      double eff = static_cast<double>(PW) * PH / (tiles * 128.0);
      if (eff > best + 1e-9) {
        best = eff;
        bbw = w;
        bbh = h;
      }
    }
  }
  *bw = bbw;
  *bh = bbh;
  *bn = 1;
  if (bbw == PW && bbh == PH) *bn = std::max(1, std::min(N, 128 / (PW * PH)));
}

static int floordiv2(int e) { return (e - (e & 1)) / 2; }

// Mirror the interior of a plane into its halo ring (nn.ReflectionPad2d of the NEXT conv's input).  One thread per
// (image, ring pixel, 16-byte channel chunk); ring pixels are enumerated as top rows, bottom rows, left / right columns.
__global__ void __launch_bounds__(256) reflect_halo_kernel(dtg_plane p, int vec_per_pix) {
  pdl_enter();
  const int hl = p.halo, Hb = p.h + 2 * hl, Wb = p.w + 2 * hl;
  const int ring = 2 * hl * Wb + 2 * hl * p.h;
  const size_t total = static_cast<size_t>(p.n) * ring * vec_per_pix;
  uint4* base = reinterpret_cast<uint4*>(p.ptr);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int v = i % vec_per_pix;
    const int r = (i / vec_per_pix) % ring;
    const int n = i / (static_cast<size_t>(vec_per_pix) * ring);
    int yb, xb;
    if (r < 2 * hl * Wb) {          // top / bottom rows
      const int row = r / Wb;
      xb = r - row * Wb;
      yb = row < hl ? row : p.h + row;          // rows [0, hl) and [h + hl, h + 2 hl)
    } else {                        // left / right columns of the interior rows
      const int q = r - 2 * hl * Wb;
      const int row = q / (2 * hl), col = q - row * (2 * hl);
      yb = row + hl;
      xb = col < hl ? col : p.w + col;
    }
    int y = yb - hl, x = xb - hl;
    y = y < 0 ? -y : (y >= p.h ? 2 * (p.h - 1) - y : y);
    x = x < 0 ? -x : (x >= p.w ? 2 * (p.w - 1) - x : x);
    const size_t img = static_cast<size_t>(n) * Hb * Wb;
    base[(img + static_cast<size_t>(yb) * Wb + xb) * vec_per_pix + v] =
        base[(img + static_cast<size_t>(y + hl) * Wb + (x + hl)) * vec_per_pix + v];
  }
}

template <bool TF32>
static int launch_igemm(const IgemmParams& p, cudaStream_t stream) {
  static int num_sms = 0;
  static bool attr_set = false;
  static std::mutex mu;
  {
    std::lock_guard<std::mutex> lk(mu);
    if (num_sms == 0) {
      int dev = 0;
      DTG_CHECK_CUDA(cudaGetDevice(&dev));
      DTG_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    if (!attr_set) {
      DTG_CHECK_CUDA(cudaFuncSetAttribute(igemm_kernel<TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      attr_set = true;
    }
  }
  const int stage_bytes = kATileBytes + p.n_umma * kRowBytes;
  const size_t smem = static_cast<size_t>(p.stages) * stage_bytes + 1024 + kBarrierBytes + 4 * kEpiWarpBytes;
  const int total = p.tiles_w * p.tiles_h * p.tiles_n * p.num_phases;
  const int grid = std::max(1, std::min(total, num_sms));
  DTG_CHECK_CUDA(launch_k(igemm_kernel<TF32>, grid, kThreads, smem, stream, p));
  return DTG_OK;
}

}  // namespace dtg

using namespace dtg;

extern "C" int dtg_conv(const dtg_conv_args* a, const dtg_plane* in, const void* w, int w_rows, int w_cols,
                        const float* bias, const dtg_plane* out, float* out_nchw, void* stream) {
  DTG_REQUIRE(a && in && w, "dtg_conv: null argument");
  {
    // out_reflect into a plane: the tiles leave through the TMA-store epilogue (which cannot mirror), then one small
    // kernel fills the halo ring -- faster than the per-thread mirrored stores of the generic epilogue
    static const bool fused_reflect = getenv("DTG_FUSED_REFLECT") != nullptr;
    if (a->out_reflect && !a->out_nchw_f32 && !fused_reflect && out && out->ptr && out->halo > 0 &&
        out->h >= out->halo + 1 && out->w >= out->halo + 1 && (out->c * elem_size(out->dtype)) % 16 == 0) {
      dtg_conv_args a2 = *a;
      a2.out_reflect = 0;
      const int rc = dtg_conv(&a2, in, w, w_rows, w_cols, bias, out, out_nchw, stream);
      if (rc != DTG_OK) return rc;
      const int vec = out->c * elem_size(out->dtype) / 16;
      const size_t total = static_cast<size_t>(out->n) * (2 * out->halo * (out->w + 2 * out->halo) + 2 * out->halo * out->h) * vec;
      DTG_CHECK_CUDA(launch_k(reflect_halo_kernel, static_cast<int>(std::max<size_t>(1, std::min<size_t>((total + 255) / 256, 148 * 16))), 256, 0, static_cast<cudaStream_t>(stream), *out, vec));
      return DTG_OK;
    }
  }
  if (a->fold_w == 2) {
    // filter column in GEMM-N (conv_tail7.cu): the weights are packed for that kernel only
    const int rc = try_launch_tail7(a, in, w, w_rows, w_cols, bias, out, out_nchw, static_cast<cudaStream_t>(stream));
    DTG_REQUIRE(rc != 1, "dtg_conv fold_w=2: geometry not eligible (stride-1 odd-kernel 'same' layer, cout <= 4, kw * cout <= 28, "
                         "input halo 0, 32/64/128 bytes per pixel, width a divisor of 128; FWD: dense fp32 NCHW output; DGRAD: "
                         "ring == pad into a 16-byte-pixel plane with halo == ring)");
    return rc;
  }
  DTG_REQUIRE(a->stride == 1 || a->stride == 2, "dtg_conv: stride %d unsupported", a->stride);
  DTG_REQUIRE(a->kh * a->kw <= kMaxTaps && a->kh >= 1 && a->kw >= 1, "dtg_conv: kernel %dx%d unsupported", a->kh, a->kw);
  DTG_REQUIRE(w_rows % 16 == 0 && w_rows >= 16 && w_rows <= 256, "dtg_conv: packed rows %d must be 16..256, multiple of 16", w_rows);
  DTG_REQUIRE(a->cout >= 1 && a->cout <= w_rows, "dtg_conv: cout %d > packed rows %d", a->cout, w_rows);
  const int es = elem_size(in->dtype);
  const bool tf32 = in->dtype == DTG_F32;
  const int KC = kRowBytes / es;
  DTG_REQUIRE((in->c * es) % 16 == 0 && (w_cols * es) % 16 == 0, "dtg_conv: channel pitch must be a multiple of 16 bytes");
  const int s = a->stride, hl = in->halo;
  const int Hb = in->h + 2 * hl, Wb = in->w + 2 * hl;

  IgemmParams p;
  memset(&p, 0, sizeof(p));
  p.N = in->n;
  p.n_umma = w_rows;
  p.e.n_umma = w_rows;
  p.kchunks = a->fold_w ? 1 : (std::min(w_cols, in->c) + KC - 1) / KC;
  // 32-byte k-steps that hold data (small-channel layers fill only part of the single 128-byte chunk)
  p.ksteps = (a->fold_w || p.kchunks > 1) ? 4 : std::max(1, (std::min(w_cols, in->c) * es + 31) / 32);
  p.e.act = a->act;
  p.e.bias = bias;
  p.e.cvalid = a->cout;
  p.e.out_reflect = a->out_reflect;
  const int OHl = a->out_h, OWl = a->out_w;  // interior extents of the output

  int PW = 0, PH = 0;
  int ntaps = 0;
  if (a->fold_w) {
    // kw-folded small-channel input: GEMM-K = (kw slot, channel) through an overlapping-row TMA view, taps = KH
    DTG_REQUIRE(s == 1 && a->ring == 0, "dtg_conv fold_w: stride 1, ring 0 only");
    DTG_REQUIRE(in->c * es == 16, "dtg_conv fold_w: input plane must hold 16 bytes per pixel (c = %d)", in->c);
    DTG_REQUIRE(a->kw <= 8 && w_cols * es == kRowBytes, "dtg_conv fold_w: kw %d > 8 slots or weights not kw-folded (cols %d)", a->kw, w_cols);
    DTG_REQUIRE(hl >= a->pad && hl >= a->kw - 1 - a->pad, "dtg_conv fold_w: needs a materialised halo (%d) >= pad", hl);
    DTG_REQUIRE(in->h + 2 * a->pad - a->kh + 1 == OHl && in->w + 2 * a->pad - a->kw + 1 == OWl, "dtg_conv fold_w: output extent mismatch");
    p.num_phases = 1;
    p.out_step = 1;
    p.ph_OH[0] = OHl;
    p.ph_OW[0] = OWl;
    p.ph_tap_begin[0] = 0;
    const bool fwd = a->mode == DTG_CONV_FWD;
    for (int kh = 0; kh < a->kh; ++kh) {
      // FWD: rows y + kh - pad, window x - pad .. ; DGRAD: rows y + pad - kh, window x + pad - (KW-1) .. (weights flipped)
      p.tap_dh[ntaps] = static_cast<short>(fwd ? kh - a->pad + hl : a->pad - kh + hl);
      p.tap_dw[ntaps] = static_cast<short>(fwd ? hl - a->pad : hl + a->pad - (a->kw - 1));
      p.tap_map[ntaps] = 0;
      p.tap_w[ntaps] = static_cast<unsigned char>(kh);
      ++ntaps;
    }
    p.ph_tap_begin[1] = ntaps;
    PW = OWl;
    PH = OHl;
  } else if (a->mode == DTG_CONV_FWD) {
    DTG_REQUIRE((in->h + 2 * a->pad - a->kh) / s + 1 == OHl && (in->w + 2 * a->pad - a->kw) / s + 1 == OWl,
                "dtg_conv fwd: output extent %dx%d inconsistent with input %dx%d k%d s%d p%d", OHl, OWl, in->h, in->w, a->kh, s, a->pad);
    DTG_REQUIRE(hl == 0 || hl >= a->pad, "dtg_conv fwd: materialised halo %d < pad %d", hl, a->pad);
    p.num_phases = 1;
    p.out_step = 1;
    p.ph_OH[0] = OHl;
    p.ph_OW[0] = OWl;
    p.ph_tap_begin[0] = 0;
    for (int kh = 0; kh < a->kh; ++kh)
      for (int kw = 0; kw < a->kw; ++kw) {
        const int eh = kh - a->pad + hl, ew = kw - a->pad + hl;
        if (s == 1) {
          p.tap_dh[ntaps] = static_cast<short>(eh);
          p.tap_dw[ntaps] = static_cast<short>(ew);
          p.tap_map[ntaps] = 0;
        } else {
          p.tap_dh[ntaps] = static_cast<short>(floordiv2(eh));
          p.tap_dw[ntaps] = static_cast<short>(floordiv2(ew));
          p.tap_map[ntaps] = static_cast<unsigned char>((eh & 1) * 2 + (ew & 1));
        }
        p.tap_w[ntaps] = static_cast<unsigned char>(kh * a->kw + kw);
        ++ntaps;
      }
    p.ph_tap_begin[1] = ntaps;
    PW = OWl;
    PH = OHl;
  } else if (a->mode == DTG_CONV_DGRAD) {
    // halo 0, or a ZERO halo ring equal to `ring` (stride 1): the flat-raster path of conv_patch2.cu then reads the ring as
    // the convolution's padding and tiles the (h + 2 ring) x (w + 2 ring) outputs of all images as one tall image
    DTG_REQUIRE(hl == 0 || (hl == a->ring && s == 1), "dtg_conv dgrad: input gradient plane must have halo 0 (or a zero halo == ring)");
    DTG_REQUIRE((OHl + 2 * a->pad - a->kh) / s + 1 == in->h && (OWl + 2 * a->pad - a->kw) / s + 1 == in->w,
                "dtg_conv dgrad: output extent %dx%d inconsistent with dy %dx%d k%d s%d p%d", OHl, OWl, in->h, in->w, a->kh, s, a->pad);
    const int R = a->ring;
    DTG_REQUIRE(R >= 0 && (s == 1 || R == 0), "dtg_conv dgrad: ring only with stride 1");
    p.out_step = s;
    if (s == 1) {
      p.num_phases = 1;
      p.ph_oh0[0] = -R;
      p.ph_ow0[0] = -R;
      p.ph_OH[0] = OHl + 2 * R;
      p.ph_OW[0] = OWl + 2 * R;
      p.ph_tap_begin[0] = 0;
      for (int kh = 0; kh < a->kh; ++kh)
        for (int kw = 0; kw < a->kw; ++kw) {
          p.tap_dh[ntaps] = static_cast<short>(a->pad - R - kh);
          p.tap_dw[ntaps] = static_cast<short>(a->pad - R - kw);
          p.tap_map[ntaps] = 0;
          p.tap_w[ntaps] = static_cast<unsigned char>(kh * a->kw + kw);
          ++ntaps;
        }
      p.ph_tap_begin[1] = ntaps;
      PW = p.ph_OW[0];
      PH = p.ph_OH[0];
    } else {
      p.num_phases = 4;
      for (int ph = 0; ph < 4; ++ph) {
        const int r = ph >> 1, c = ph & 1;
        p.ph_oh0[ph] = r;
        p.ph_ow0[ph] = c;
        p.ph_OH[ph] = (OHl - r + 1) / 2;
        p.ph_OW[ph] = (OWl - c + 1) / 2;
        p.ph_tap_begin[ph] = ntaps;
        for (int kh = 0; kh < a->kh; ++kh) {
          if (((r + a->pad - kh) & 1) != 0) continue;
          for (int kw = 0; kw < a->kw; ++kw) {
            if (((c + a->pad - kw) & 1) != 0) continue;
            p.tap_dh[ntaps] = static_cast<short>((r + a->pad - kh) / 2);
            p.tap_dw[ntaps] = static_cast<short>((c + a->pad - kw) / 2);
            p.tap_map[ntaps] = 0;
            p.tap_w[ntaps] = static_cast<unsigned char>(kh * a->kw + kw);
            ++ntaps;
          }
        }
        DTG_REQUIRE(ntaps > p.ph_tap_begin[ph], "dtg_conv dgrad: empty parity phase (k%d s2)", a->kh);
        PW = std::max(PW, p.ph_OW[ph]);
        PH = std::max(PH, p.ph_OH[ph]);
      }
      p.ph_tap_begin[4] = ntaps;
    }
  } else {
    DTG_REQUIRE(false, "dtg_conv: bad mode %d", a->mode);
  }

  choose_patch(PW, PH, in->n, &p.bw, &p.bh, &p.bn);
  p.tiles_w = (PW + p.bw - 1) / p.bw;
  p.tiles_h = (PH + p.bh - 1) / p.bh;
  p.tiles_n = (in->n + p.bn - 1) / p.bn;

  // output
  if (a->out_nchw_f32) {
    DTG_REQUIRE(out_nchw != nullptr, "dtg_conv: out_nchw is null");
    p.e.out = out_nchw;
    p.e.out_nchw = 1;
    p.e.out_H = OHl;
    p.e.out_W = OWl;
    p.e.out_halo = 0;
    p.e.out_C = a->cout;
    DTG_REQUIRE(a->ring == 0, "dtg_conv: ring with NCHW output");
  } else {
    DTG_REQUIRE(out != nullptr && out->ptr != nullptr, "dtg_conv: out plane is null");
    DTG_REQUIRE(out->dtype == in->dtype, "dtg_conv: out dtype must equal in dtype");
    DTG_REQUIRE(out->h == OHl && out->w == OWl && out->n == in->n, "dtg_conv: out plane extent mismatch");
    DTG_REQUIRE(out->c % (16 / es) == 0 && out->c <= w_rows, "dtg_conv: out plane channels %d vs packed rows %d", out->c, w_rows);
    DTG_REQUIRE(out->halo >= (a->mode == DTG_CONV_DGRAD ? a->ring : 0), "dtg_conv: out halo < ring");
    p.e.out = out->ptr;
    p.e.out_C = out->c;
    p.e.out_halo = out->halo;
    p.e.out_H = out->h;
    p.e.out_W = out->w;
  }

  const int stage_bytes = kATileBytes + p.n_umma * kRowBytes;
  p.stages = std::max(2, std::min(8, (tensor_smem_budget() - 31 * 1024) / stage_bytes));    // + barriers, epilogue staging, alignment
  int cols = 32;
  while (cols < 2 * p.n_umma) cols <<= 1;
  p.tmem_cols = cols;
  DTG_REQUIRE(a->act != DTG_ACT_TANH || a->out_nchw_f32, "dtg_conv: tanh is only fused into dense NCHW head outputs");
  if (a->out_reflect && !a->out_nchw_f32)
    DTG_REQUIRE(p.e.out_H >= 2 * p.e.out_halo + 2 && p.e.out_W >= 2 * p.e.out_halo + 2, "dtg_conv: reflect halo %d too wide for %dx%d", p.e.out_halo, p.e.out_H, p.e.out_W);
  if (a->mode == DTG_CONV_DGRAD && hl > 0 && !a->fold_w) {
    p.flat_dgrad = 1;
    const int rc2 = try_launch_pconv2(p, in, w, w_rows, w_cols, a->kh * a->kw, static_cast<cudaStream_t>(stream));
    DTG_REQUIRE(rc2 != 1, "dtg_conv dgrad: a haloed dy plane needs the flat-raster path (full 128-byte channel chunks, <= 128 "
                          "output channels, output plane with the same halo, n * (h + 2 halo) * (w + 2 halo) divisible by 8)");
    return rc2;
  }
  {
    const int rc = try_launch_pconv(p, in, w, w_rows, w_cols, a->fold_w ? a->kh : a->kh * a->kw, a->fold_w,
                                    static_cast<cudaStream_t>(stream));
    if (rc <= 0) return rc;   // launched (0) or failed (<0); 1 = not eligible
    DTG_REQUIRE(!a->fold_w, "dtg_conv fold_w: geometry not supported by the patch kernel");
    const int rc2 = try_launch_pconv2(p, in, w, w_rows, w_cols, a->kh * a->kw, static_cast<cudaStream_t>(stream));
    if (rc2 <= 0) return rc2;
  }

  // activation tensor maps
  const bool fwd_s2 = (a->mode == DTG_CONV_FWD && s == 2);
  for (int m = 0; m < 4; ++m) {
    const int ph = fwd_s2 ? (m >> 1) : 0, pw = fwd_s2 ? (m & 1) : 0;
    const int st = fwd_s2 ? 2 : 1;
    uint64_t dims[4] = {static_cast<uint64_t>(in->c), static_cast<uint64_t>(std::max(1, (Wb - pw + st - 1) / st)),
                        static_cast<uint64_t>(std::max(1, (Hb - ph + st - 1) / st)), static_cast<uint64_t>(in->n)};
    uint64_t strides[3] = {static_cast<uint64_t>(st) * in->c * es, static_cast<uint64_t>(st) * Wb * in->c * es,
                           static_cast<uint64_t>(Hb) * Wb * in->c * es};
    uint32_t box[4] = {static_cast<uint32_t>(KC), static_cast<uint32_t>(p.bw), static_cast<uint32_t>(p.bh),
                       static_cast<uint32_t>(p.bn)};
    uint8_t* base = reinterpret_cast<uint8_t*>(in->ptr) + (static_cast<size_t>(ph) * Wb + pw) * in->c * es;
    int rc = encode_tiled(&p.tmA[m], in->dtype, 4, base, dims, strides, box, 1);
    if (rc != DTG_OK) return rc;
    if (!fwd_s2) {
      for (int k = 1; k < 4; ++k) p.tmA[k] = p.tmA[0];
      break;
    }
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(w_cols), static_cast<uint64_t>(w_rows) * a->kh * a->kw};
    uint64_t strides[1] = {static_cast<uint64_t>(w_cols) * es};
    uint32_t box[2] = {static_cast<uint32_t>(KC), static_cast<uint32_t>(w_rows)};
    int rc = encode_tiled(&p.tmB, in->dtype, 2, const_cast<void*>(w), dims, strides, box, 1);
    if (rc != DTG_OK) return rc;
  }

  return tf32 ? launch_igemm<true>(p, static_cast<cudaStream_t>(stream))
              : launch_igemm<false>(p, static_cast<cudaStream_t>(stream));
}
