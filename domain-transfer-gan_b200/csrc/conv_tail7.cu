// The generators' 7x7 tail (networks.py:187-188 / 242-243: Conv2d(ngf, output_nc, 7, padding=3) + Tanh) for sm_100a with
// the filter COLUMN folded into GEMM-N.
//
// With cout = 3 the ordinary mapping (GEMM-N = cout padded to 16, one tap per MMA) issues 49 taps x cin/16 k-steps =
// 98 tcgen05.mma of ~36 clk per 128-pixel tile: the tensor pipe is occupied by operand fetch, not math
// (tools/umma_rate_test.cu), and the layer ran at ~50 TFLOP/s.  Here
//     D[p][(kw, co)] = sum_{kh, ci} x[py + kh - 3][px][ci] * w[co][ci][kh][kw]          (N = 7 * cout <= 32)
// needs only the 7 filter ROWS as taps -- 14 MMAs of N = 32 per tile for cin = 32 -- and every tap is the same
// full-width patch rows shifted by kh image rows (the patch has no x halo, so 128 consecutive GEMM rows are 128
// consecutive pixels of 128 / W image rows and the 8-row group stride is uniform).  The epilogue adds the seven column
// partials with their pixel shift,  y[oy][ox][co] = b[co] + sum_kw D[(oy, ox + kw - 3)][(kw, co)]  (terms outside the row
// are the conv's zero padding), through a bank-conflict-free fp32 staging tile in shared memory, applies tanh and writes
// the dense fp32 NCHW head.  Rows outside the image are TMA zero fill.
// Warp roles as in conv_patch.cu: warp0 TMA producer, warp1 MMA issuer (+TMEM alloc), warps2-5 epilogue.
#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "conv_epilogue.cuh"

namespace dtg {

constexpr int kT7MaxMma = 32;
constexpr int kT7Acc = 4;           // TMEM accumulator buffers
constexpr int kT7N = 32;            // GEMM-N (kw, cout) padded
constexpr int kT7Pitch = 29;        // fp32 staging row pitch: odd -> the shifted reads of a warp hit 32 different banks

struct Tail7Params {
  CUtensorMap tmA, tmB;
  int H, W, N, cout, kw, kh;
  int rpt;               // image rows per 128-pixel tile (128 / W)
  int PH;                // patch rows = rpt + kh - 1
  int pad;
  int tiles_per_img;
  int a_stage_bytes, a_stages, rb, layout, b_tap_bytes, rb_elems;
  int nmma;
  unsigned short a_off[kT7MaxMma], b_off[kT7MaxMma];
  int act;
  int full;              // 1: data gradient of a 7x7 HEAD w.r.t. its padded input ("full" correlation: (H+kh-1) x (W+kw-1)
                         //    outputs, taps and column shifts mirrored), written as 16-byte pixels of a haloed plane
  int row_org;           // first patch row relative to the tile's first output row: -pad (forward) / -(kh-1) (full)
  int OH, OW;            // output rows / columns per image
  dtg_plane outp;        // full: destination plane (16 bytes per pixel, halo = pad)
  int dbg;               // DTG_T7_DBG experiments: 1 = one MMA per tile, 2 = skip the patch loads, 4 = skip the epilogue body
  const float* bias;
  float* out;            // [N][cout][H][W] fp32
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <bool TF32>
__global__ void __launch_bounds__(kThreads, 1) tail7_kernel(const __grid_constant__ Tail7Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  pdl_trigger();
  const int S = p.a_stages;
  uint8_t* sA = smem;
  uint8_t* sB = smem + S * p.a_stage_bytes;
  const int b_bytes = p.kh * p.b_tap_bytes;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(sB + ((b_bytes + 1023) & ~1023));
  uint64_t* bar_empty = bar_full + S;
  uint64_t* bar_tfull = bar_empty + S;
  uint64_t* bar_tempty = bar_tfull + kT7Acc;
  uint64_t* bar_b = bar_tempty + kT7Acc;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_b + 1);
  float* stage_f = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bar_full) + 1024);   // 2 x [128][kT7Pitch] fp32

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < S; ++i) {
        mbar_init(&bar_full[i], 1);
        mbar_init(&bar_empty[i], 1);
      }
      for (int i = 0; i < kT7Acc; ++i) {
        mbar_init(&bar_tfull[i], 1);
        mbar_init(&bar_tempty[i], 4);
      }
      mbar_init(bar_b, 1);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, kT7Acc * kT7N);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  const int total_tiles = p.tiles_per_img * p.N;
  const uint32_t a_tx = static_cast<uint32_t>(p.PH) * p.W * p.rb;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) mbar_expect_tx(bar_b, static_cast<uint32_t>(b_bytes));
    __syncwarp();
    for (int t = 0; t < p.kh; ++t)
      if (elect_one()) tma_load_2d(sB + t * p.b_tap_bytes, &p.tmB, bar_b, 0, t * kT7N);
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n = tile / p.tiles_per_img;
      const int th = tile - n * p.tiles_per_img;
      mbar_wait(&bar_empty[stage], phase ^ 1);
      if (p.dbg & 2) {
        if (elect_one()) mbar_arrive(&bar_full[stage]);
      } else if (elect_one()) {
        mbar_expect_tx(&bar_full[stage], a_tx);
        tma_load_4d(sA + stage * p.a_stage_bytes, &p.tmA, &bar_full[stage], 0, 0, th * p.rpt + p.row_org, n);
      }
      __syncwarp();
      if (++stage == S) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform, elect-predicated: see conv_patch.cu) =====================
    const uint32_t idesc = umma_idesc(TF32 ? 2u : 1u, 0u, 0u, kTileM, kT7N);
    const uint32_t hi = ((8u * p.rb) >> 4) | (1u << 14) | (static_cast<uint32_t>(p.layout) << 29);
    mbar_wait(bar_b, 0);
    tc_fence_after();
    const uint32_t b_lo0 = ((smem_u32(sB) >> 4) & 0x3FFFu) | (1u << 16);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int buf = it & (kT7Acc - 1);
      const uint32_t use = static_cast<uint32_t>(it / kT7Acc);
      mbar_wait(&bar_tempty[buf], (use & 1) ^ 1);
      mbar_wait(&bar_full[stage], phase);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * kT7N;
      const uint32_t a_lo0 = ((smem_u32(sA + stage * p.a_stage_bytes) >> 4) & 0x3FFFu) | (1u << 16);
      if (elect_one()) {
        uint32_t acc = 0u;
        for (int i = 0; i < p.nmma; ++i) {
          tc_mma<TF32>(d_tmem, (static_cast<uint64_t>(hi) << 32) | (a_lo0 + p.a_off[i]),
                       (static_cast<uint64_t>(hi) << 32) | (b_lo0 + p.b_off[i]), idesc, acc);
          acc = 1u;
        }
        tc_commit(&bar_empty[stage]);
        tc_commit(&bar_tfull[buf]);
      }
      __syncwarp();
      if (++stage == S) {
        stage = 0;
        phase ^= 1;
      }
      ++it;
    }
  } else {
    // ===================== epilogue: shift-add of the kw partials =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;            // GEMM row = pixel of the tile, raster order over rpt image rows
    const int r = row / p.W, ox = row - r * p.W;
    const float b0 = p.bias ? p.bias[0] : 0.f;
    const float b1 = (p.bias && p.cout > 1) ? p.bias[1] : 0.f;
    const float b2 = (p.bias && p.cout > 2) ? p.bias[2] : 0.f;
    const float b3 = (p.bias && p.cout > 3) ? p.bias[3] : 0.f;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n = tile / p.tiles_per_img;
      const int th = tile - n * p.tiles_per_img;
      const int buf = it & (kT7Acc - 1);
      const uint32_t use = static_cast<uint32_t>(it / kT7Acc);
      float* st = stage_f + (it & 1) * (kTileM * kT7Pitch);
      mbar_wait(&bar_tfull[buf], use & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + buf * kT7N;
      if (p.dbg & 4) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_tempty[buf]);
        ++it;
        continue;
      }
      uint32_t v0[16], v1[16];
      tmem_ld16(taddr, v0);
      tmem_ld16(taddr + 16, v1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_tempty[buf]);       // the accumulator is in registers: release it to the MMA warp
      float* mine = st + row * kT7Pitch;
#pragma unroll
      for (int j = 0; j < 16; ++j) mine[j] = __uint_as_float(v0[j]);
#pragma unroll
      for (int j = 0; j < 12; ++j) mine[16 + j] = __uint_as_float(v1[j]);
      named_bar_sync(1, 128);
      const int oy = th * p.rpt + r;
      if (oy < p.OH) {
        // registers only: constant trip counts, runtime bounds as predicates (runtime-indexed arrays would go to local
        // memory: 38 us instead of 14 for 32 -> 3 channels at 64x64 x 80).  A staging row holds 29 floats, so reading four
        // columns from kw * cout is always in bounds; columns >= cout are simply not stored.
        const float* srow = st + (r * p.W) * kT7Pitch;
        // forward: y[ox] = sum_kw D[ox + kw - pad][kw];   full: dx[o] = sum_kw D[o - kw][kw]  (o = ox, and o = W + ox for the
        // kw - 1 extra columns of the wider output)
        for (int o = ox; o < p.OW; o += p.W) {
          float a0 = b0, a1 = b1, a2 = b2, a3 = b3;
#pragma unroll
          for (int kw = 0; kw < 8; ++kw) {
            const int sx = p.full ? o - kw : o + kw - p.pad;
            if (kw < p.kw && sx >= 0 && sx < p.W) {
              const float* src = srow + sx * kT7Pitch + kw * p.cout;
              a0 += src[0];
              a1 += src[1];
              a2 += src[2];
              a3 += src[3];
            }
          }
          if (!p.full) {
            const size_t plane = static_cast<size_t>(p.H) * p.W;
            float* op = p.out + (static_cast<size_t>(n) * p.cout * p.H + oy) * p.W + o;
            op[0] = apply_act(a0, p.act);
            if (p.cout > 1) op[plane] = apply_act(a1, p.act);
            if (p.cout > 2) op[2 * plane] = apply_act(a2, p.act);
            if (p.cout > 3) op[3 * plane] = apply_act(a3, p.act);
          } else {
            if (p.cout < 2) a1 = 0.f;
            if (p.cout < 3) a2 = 0.f;
            if (p.cout < 4) a3 = 0.f;
            const size_t pix = (static_cast<size_t>(n) * p.OH + oy) * p.OW + o;
            if (p.outp.dtype == DTG_BF16) {
              __nv_bfloat162 lo = __floats2bfloat162_rn(a0, a1), hi = __floats2bfloat162_rn(a2, a3);
              uint4 v = make_uint4(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi), 0u, 0u);
              reinterpret_cast<uint4*>(p.outp.ptr)[pix] = v;
            } else {
              reinterpret_cast<float4*>(p.outp.ptr)[pix] = make_float4(round_tf32(a0), round_tf32(a1), round_tf32(a2), round_tf32(a3));
            }
          }
        }
      }
      ++it;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kT7Acc * kT7N);
  }
}

// returns DTG_OK after launching, 1 if the geometry is not eligible (the caller reports an error: the weights are packed
// for this kernel only)
int try_launch_tail7(const dtg_conv_args* a, const dtg_plane* in, const void* w, int w_rows, int w_cols, const float* bias,
                     const dtg_plane* out, float* out_nchw, cudaStream_t stream) {
  const int es = elem_size(in->dtype);
  const bool tf32 = in->dtype == DTG_F32;
  const int rb = in->c * es;
  const bool full = a->mode == DTG_CONV_DGRAD;
  if (a->stride != 1 || in->halo != 0) return 1;
  if (!full && !out_nchw) return 1;
  // full: gradient w.r.t. the padded input of a "same" head (ring == pad): every pixel of the haloed 16-byte-pixel plane
  if (full && (a->ring != a->pad || !out || !out->ptr || out->halo != a->ring || out->c * es != 16 || out->dtype != in->dtype ||
               out->n != in->n || out->h != in->h || out->w != in->w || bias != nullptr || a->act != DTG_ACT_NONE))
    return 1;
  if (rb != 32 && rb != 64 && rb != 128) return 1;
  if (a->cout < 1 || a->cout > 4 || a->kw * a->cout > 28 || a->kh > 8 || w_rows != kT7N || w_cols != in->c) return 1;
  if (a->kh != 2 * a->pad + 1 || a->kw != 2 * a->pad + 1 || a->out_h != in->h || a->out_w != in->w) return 1;
  const int W = in->w, H = in->h;
  if (W < 8 || W > 128 || (128 % W) != 0) return 1;
  Tail7Params p;
  memset(&p, 0, sizeof(p));
  p.H = H;
  p.W = W;
  p.N = in->n;
  p.cout = a->cout;
  p.kw = a->kw;
  p.kh = a->kh;
  p.pad = a->pad;
  p.rpt = 128 / W;
  p.PH = p.rpt + a->kh - 1;
  p.full = full ? 1 : 0;
  p.row_org = full ? -(a->kh - 1) : -a->pad;
  p.OH = full ? H + a->kh - 1 : H;
  p.OW = full ? W + a->kw - 1 : W;
  if (p.OW > 2 * W) return 1;
  if (full) p.outp = *out;
  p.tiles_per_img = (p.OH + p.rpt - 1) / p.rpt;
  p.rb = rb;
  p.layout = rb == 128 ? 2 : (rb == 64 ? 4 : 6);
  p.rb_elems = rb / es;
  p.b_tap_bytes = kT7N * rb;
  p.a_stage_bytes = (p.PH * W * rb + 1023) & ~1023;
  const int ks = rb / 32;
  p.nmma = a->kh * ks;
  if (p.nmma > kT7MaxMma) return 1;
  for (int t = 0; t < a->kh; ++t)
    for (int j = 0; j < ks; ++j) {
      p.a_off[t * ks + j] = static_cast<unsigned short>((t * W * rb + j * 32) >> 4);
      // full: patch row t holds dy row (oy - (kh-1) + t), which meets filter row kh-1-t
      p.b_off[t * ks + j] = static_cast<unsigned short>(((full ? a->kh - 1 - t : t) * p.b_tap_bytes + j * 32) >> 4);
    }
  const int b_bytes = a->kh * p.b_tap_bytes;
  const int fixed = 1024 + ((b_bytes + 1023) & ~1023) + 1024 + 2 * kTileM * kT7Pitch * 4;
  const int budget = tensor_smem_budget() - fixed;
  if (budget < 2 * p.a_stage_bytes) return 1;
  p.a_stages = std::max(2, std::min(4, budget / p.a_stage_bytes));
  p.act = a->act;
  {
    const char* d = getenv("DTG_T7_DBG");
    p.dbg = d ? atoi(d) : 0;
    if (p.dbg & 1) p.nmma = 1;
  }
  p.bias = bias;
  p.out = out_nchw;
  {
    uint64_t dims[4] = {static_cast<uint64_t>(in->c), static_cast<uint64_t>(W), static_cast<uint64_t>(H), static_cast<uint64_t>(in->n)};
    uint64_t strides[3] = {static_cast<uint64_t>(rb), static_cast<uint64_t>(W) * rb, static_cast<uint64_t>(H) * W * rb};
    uint32_t box[4] = {static_cast<uint32_t>(in->c), static_cast<uint32_t>(W), static_cast<uint32_t>(p.PH), 1u};
    int rc = encode_tiled(&p.tmA, in->dtype, 4, in->ptr, dims, strides, box, rb == 128 ? 1 : (rb == 64 ? 3 : 4));
    if (rc != DTG_OK) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(w_cols), static_cast<uint64_t>(kT7N) * a->kh};
    uint64_t strides[1] = {static_cast<uint64_t>(w_cols) * es};
    uint32_t box[2] = {static_cast<uint32_t>(rb / es), static_cast<uint32_t>(kT7N)};
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
        DTG_CHECK_CUDA(cudaFuncSetAttribute(tail7_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      else
        DTG_CHECK_CUDA(cudaFuncSetAttribute(tail7_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      attr_set[tf32 ? 1 : 0] = true;
    }
  }
  const size_t smem = static_cast<size_t>(fixed) + static_cast<size_t>(p.a_stages) * p.a_stage_bytes;
  const int total = p.tiles_per_img * p.N;
  const int grid = std::max(1, std::min(total, num_sms));
  if (tf32)
    DTG_CHECK_CUDA(launch_k(tail7_kernel<true>, grid, kThreads, smem, stream, p));
  else
    DTG_CHECK_CUDA(launch_k(tail7_kernel<false>, grid, kThreads, smem, stream, p));
  return DTG_OK;
}

}  // namespace dtg
