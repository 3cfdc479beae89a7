// Shared epilogue of the tcgen05 convolution kernels: TMEM accumulator rows -> registers -> bias /
// activation / convert -> per-warp smem staging (128-byte channel chunks) -> global rows written as full
// 128-byte lines (8 lanes x 16 B per row, 4 rows per warp instruction).  Mirrored copies into the output
// plane's reflect halo go through the same coalesced path (up to 4 destinations per row); heads
// (<= 16 channels) go straight to a dense fp32 NCHW tensor.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace dtg {

constexpr int kEpiPitch = 144;                              // 128-byte chunk row + 16 B pad (bank spread)
constexpr int kEpiDst = 4;                                  // destinations per row: interior + reflect mirrors
constexpr int kEpiWarpBytes = 32 * kEpiPitch + 32 * kEpiDst * 4;   // staging rows + destination-offset table

struct EpiParams {
  CUtensorMap tmOut;   // use_tma: NHWC view of exactly the valid output region {C, OW, OH, N}, box {row_elems, 8, 4, 1}
  int use_tma;         // 1: rows leave through per-warp TMA stores (epilogue_tma), which clip ragged tiles
  int row_bytes;       // use_tma: bytes per pixel per channel chunk (32 / 64 / 128), = smem staging row and swizzle span
  void* out;
  int out_nchw;  // 1: dense fp32 NCHW [N][cvalid][out_H][out_W]
  int out_C, out_halo, out_H, out_W;
  int cvalid;
  int act;
  int out_reflect;
  int n_umma;
  const float* bias;
};

constexpr int kMaxTaps = 64;
constexpr int kRowBytes = 128;
constexpr int kTileM = 128;
constexpr int kATileBytes = kTileM * kRowBytes;  // 16 KB
constexpr int kThreads = 192;
constexpr int kBarrierBytes = 256;

struct IgemmParams {
  CUtensorMap tmA[4];
  CUtensorMap tmB;
  int bw, bh, bn;
  int tiles_w, tiles_h, tiles_n;
  int num_phases;
  int ph_tap_begin[5];
  int ph_oh0[4], ph_ow0[4];
  int ph_OH[4], ph_OW[4];
  int out_step;
  int N;
  short tap_dh[kMaxTaps], tap_dw[kMaxTaps];
  unsigned char tap_map[kMaxTaps], tap_w[kMaxTaps];
  int kchunks;
  int ksteps;  // tcgen05.mma per stage (4 = full 128-byte chunk)
  int n_umma;  // = packed weight rows per tap
  int stages;
  int tmem_cols;
  int flat_dgrad;   // DGRAD whose dy plane carries a zero halo == ring: conv_patch2.cu's flat-raster mode (set by dtg_conv)
  EpiParams e;
};


// Patch-resident variant (conv_patch.cu): returns DTG_OK after launching, or 1 if the geometry is not eligible
// (the caller then falls back to the per-tap igemm kernel).
int try_launch_pconv(const IgemmParams& p, const dtg_plane* in, const void* w, int w_rows, int w_cols, int taps_total,
                     int fold_w, cudaStream_t stream);

// Filter-column-in-GEMM-N tail (conv_tail7.cu, dtg_conv_args.fold_w == 2): DTG_OK after launching, 1 if not eligible.
int try_launch_tail7(const dtg_conv_args* a, const dtg_plane* in, const void* w, int w_rows, int w_cols, const float* bias,
                     const dtg_plane* out, float* out_nchw, cudaStream_t stream);

// ---- TMA-store epilogue ------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

constexpr int kEpiTmaBuf = 4096;                     // 32 rows x 128 B staging buffer
constexpr int kEpiTmaWarpBytes = 2 * kEpiTmaBuf;     // double-buffered per warp (1024-byte aligned)

template <bool TF32, int NV>
__device__ __forceinline__ void epi_stage_cols(const EpiParams& p, uint32_t (&v)[NV], int col0, uint32_t srow,
                                               uint32_t swz_mask, float slope) {
  // v: NV fp32 accumulator columns [col0, col0+NV) of this thread's row; srow: smem address of the row start.
  // activation = max(x, slope * x): slope 1 (none), 0 (ReLU), 0.2 (LeakyReLU) -- branch-free
  constexpr int ES = TF32 ? 4 : 2;
  if (p.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < NV; ++j)
      if (col0 + j < p.cvalid) v[j] = __float_as_uint(__uint_as_float(v[j]) + __ldg(p.bias + col0 + j));
  }
  if (slope != 1.f) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const float x = __uint_as_float(v[j]);
      v[j] = __float_as_uint(fmaxf(x, slope * x));
    }
  }
  const uint32_t boff = static_cast<uint32_t>(col0 % (128 / ES)) * ES;     // byte offset inside the chunk row
  if constexpr (!TF32) {
#pragma unroll
    for (int q = 0; q < NV / 8; ++q) {
      uint32_t pk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(v[8 * q + 2 * j]), __uint_as_float(v[8 * q + 2 * j + 1]));
        pk[j] = *reinterpret_cast<uint32_t*>(&h2);
      }
      const uint32_t a = srow + boff + 16 * q;
      st_shared_v4(a ^ (((a >> 7) & swz_mask) << 4), pk[0], pk[1], pk[2], pk[3]);
    }
  } else {
#pragma unroll
    for (int q = 0; q < NV / 4; ++q) {
      const uint32_t a = srow + boff + 16 * q;
      st_shared_v4(a ^ (((a >> 7) & swz_mask) << 4), __float_as_uint(round_tf32(__uint_as_float(v[4 * q]))),
                   __float_as_uint(round_tf32(__uint_as_float(v[4 * q + 1]))),
                   __float_as_uint(round_tf32(__uint_as_float(v[4 * q + 2]))),
                   __float_as_uint(round_tf32(__uint_as_float(v[4 * q + 3]))));
    }
  }
}

// One warp drains its 32 accumulator rows = a box of 8 x 4 pixels at (c1, c2, c3) of the output view, chunk by
// chunk (128 bytes of channels): TMEM -> registers -> bias/activation/convert -> swizzled smem rows -> one TMA
// store per chunk (clipped to the tensor extent, so ragged tiles need no masks).  `sbase`: this warp's
// kEpiTmaWarpBytes staging area (1024-byte aligned); `cnt`: running count of this warp's stores (buffer parity).
template <bool TF32>
__device__ __forceinline__ void epilogue_tma(const EpiParams& p, uint32_t taddr, int c1, int c2, int c3, uint8_t* sbase,
                                             uint32_t& cnt, int lane, int dbg = 0) {
  constexpr int ES = TF32 ? 4 : 2;
  constexpr int CH = 128 / ES;
  const int ncol = p.row_bytes / ES;                 // columns per chunk
  const int nchunks = (p.out_C * ES + 127) / 128;
  const uint32_t swz_mask = p.row_bytes == 128 ? 7u : (p.row_bytes == 64 ? 3u : 1u);
  const float slope = p.act == DTG_ACT_RELU ? 0.f : (p.act == DTG_ACT_LRELU ? 0.2f : 1.f);   // tanh: legacy path only
  for (int ch = 0; ch < nchunks; ++ch) {
    uint8_t* sb = sbase + (cnt & 1u) * kEpiTmaBuf;
    if (cnt >= 2 && !(dbg & 64)) {
      if (elect_one()) bulk_wait_read<1>();          // the store issued two chunks ago has finished reading `sb`
      __syncwarp();
    }
    const uint32_t srow = smem_u32(sb) + lane * p.row_bytes;
    if (dbg & 32) {
    } else if (ncol >= 32) {
      for (int c0 = 0; c0 < ncol; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + ch * CH + c0, v);
        tmem_ld_wait();
        epi_stage_cols<TF32, 32>(p, v, ch * CH + c0, srow, swz_mask, slope);
      }
    } else {
      uint32_t v[16];
      tmem_ld16(taddr + ch * CH, v);
      tmem_ld_wait();
      epi_stage_cols<TF32, 16>(p, v, ch * CH, srow, swz_mask, slope);
    }
    fence_async_smem();
    __syncwarp();
    if (!(dbg & 16) && elect_one()) {
      tma_store_4d(&p.tmOut, sb, ch * CH, c1, c2, c3);
      bulk_commit();
    }
    ++cnt;
  }
}

// Streamed-weight, two-tiles-per-item variant for 128-channel 3x3 layers (conv_patch2.cu); same return convention.
int try_launch_pconv2(const IgemmParams& p, const dtg_plane* in, const void* w, int w_rows, int w_cols, int taps_total,
                      cudaStream_t stream);

// One warp (TMEM lane quadrant) drains its 32 accumulator rows.  Row `lane` is output pixel (n, oh, ow)
// (valid = inside the output); taddr addresses the quadrant's lanes at the accumulator's first column.
// Call epilogue_prepare() (fills the destination table) before waiting for the accumulator, then
// epilogue_rows().
template <bool TF32>
__device__ __forceinline__ void epilogue_prepare(const EpiParams& p, bool valid, int n, int oh, int ow, uint8_t* stile,
                                                 int lane) {
  constexpr int ES = TF32 ? 4 : 2;
  if (p.out_nchw) return;
  uint32_t* soff = reinterpret_cast<uint32_t*>(stile + 32 * kEpiPitch);   // [32][kEpiDst] offsets in 16-byte units
  const int Hb = p.out_H + 2 * p.out_halo, Wb = p.out_W + 2 * p.out_halo;
  uint32_t d[kEpiDst] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
  if (valid) {
    int hts[3], wts[3];
    int nh = 1, nw = 1;
    hts[0] = oh;
    wts[0] = ow;
    if (p.out_reflect) {
      nh = reflect_targets(oh, p.out_H, p.out_halo, hts);
      nw = reflect_targets(ow, p.out_W, p.out_halo, wts);
    }
    int k = 0;
    for (int i = 0; i < nh; ++i)
      for (int j = 0; j < nw; ++j) {
        if (k < kEpiDst) {
          const size_t pix = (static_cast<size_t>(n) * Hb + (hts[i] + p.out_halo)) * Wb + (wts[j] + p.out_halo);
          d[k] = static_cast<uint32_t>((pix * p.out_C * ES) >> 4);
        }
        ++k;
      }
  }
  *reinterpret_cast<uint4*>(soff + lane * kEpiDst) = make_uint4(d[0], d[1], d[2], d[3]);
}

template <bool TF32>
__device__ __forceinline__ void epilogue_rows(const EpiParams& p, uint32_t taddr, bool valid, int n, int oh, int ow,
                                              uint8_t* stile, int lane) {
  using OutT = typename std::conditional<TF32, float, __nv_bfloat16>::type;
  constexpr int ES = sizeof(OutT);
  constexpr int CHUNK_CH = 128 / ES;   // channels per 128-byte chunk
  constexpr int PIECE_CH = 16 / ES;    // channels per 16-byte piece
  const bool has_bias = p.bias != nullptr;
  if (p.out_nchw) {
    // heads (<= 16 channels): lanes are consecutive pixels, so per-channel stores are already coalesced
    for (int c0 = 0; c0 < p.n_umma; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(taddr + c0, v);
      tmem_ld_wait();
      if (!valid) continue;
      float* o = reinterpret_cast<float*>(p.out);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int c = c0 + j;
        if (c < p.cvalid) {
          float x = __uint_as_float(v[j]);
          if (has_bias) x += __ldg(p.bias + c);
          o[((static_cast<size_t>(n) * p.cvalid + c) * p.out_H + oh) * p.out_W + ow] = apply_act(x, p.act);
        }
      }
    }
    return;
  }
  const uint32_t* soff = reinterpret_cast<const uint32_t*>(stile + 32 * kEpiPitch);
  const float slope = p.act == DTG_ACT_RELU ? 0.f : (p.act == DTG_ACT_LRELU ? 0.2f : 1.f);
  const int ncols = min(p.n_umma, p.out_C);
  const int piece = lane & 7, rsub = lane >> 3;
  uint8_t* const outp = reinterpret_cast<uint8_t*>(p.out);
  const int ndst = p.out_reflect ? kEpiDst : 1;
  for (int cbase = 0; cbase < ncols; cbase += CHUNK_CH) {
    __syncwarp();
    const int cend = min(cbase + CHUNK_CH, p.n_umma);
    for (int c0 = cbase; c0 < cend; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(taddr + c0, v);
      tmem_ld_wait();
      if (has_bias) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c0 + j < p.cvalid) v[j] = __float_as_uint(__uint_as_float(v[j]) + __ldg(p.bias + c0 + j));
      }
      if (slope != 1.f) {     // max(x, slope*x): ReLU (0) / LeakyReLU (0.2), branch-free (tanh: NCHW heads only)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float x = __uint_as_float(v[j]);
          v[j] = __float_as_uint(fmaxf(x, slope * x));
        }
      }
      uint8_t* dst = stile + lane * kEpiPitch + (c0 - cbase) * ES;
      if constexpr (!TF32) {
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
          pk[j] = *reinterpret_cast<uint32_t*>(&h2);
        }
        reinterpret_cast<uint4*>(dst)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        reinterpret_cast<uint4*>(dst)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          reinterpret_cast<float4*>(dst)[q] =
              make_float4(round_tf32(__uint_as_float(v[4 * q])), round_tf32(__uint_as_float(v[4 * q + 1])),
                          round_tf32(__uint_as_float(v[4 * q + 2])), round_tf32(__uint_as_float(v[4 * q + 3])));
      }
    }
    __syncwarp();
    const int ch_of_piece = cbase + piece * PIECE_CH;
    if (ch_of_piece < p.out_C) {
      uint8_t* const obase = outp + static_cast<size_t>(ch_of_piece) * ES;
#pragma unroll
      for (int r4 = 0; r4 < 32; r4 += 4) {
        const int rr = r4 + rsub;
        const uint4 val = *reinterpret_cast<const uint4*>(stile + rr * kEpiPitch + piece * 16);
        for (int d = 0; d < ndst; ++d) {
          const uint32_t off = soff[rr * kEpiDst + d];
          if (off == 0xFFFFFFFFu) break;
          *reinterpret_cast<uint4*>(obase + (static_cast<size_t>(off) << 4)) = val;
        }
      }
    }
  }
  __syncwarp();
}

}  // namespace dtg
