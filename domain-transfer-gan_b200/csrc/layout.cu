// Layout kernels: weight packing for the implicit-GEMM operands, NCHW fp32 <-> NHWC plane
// conversion at network entry/exit (with reflection halo), gradient fan-in gather (+tanh backward),
// per-channel sums (bias gradients).  All HBM-bound, coalesced along the contiguous axis.
#include <algorithm>

#include "common.cuh"

namespace dtg {

template <typename T>
__device__ __forceinline__ T to_elem(float v);
template <>
__device__ __forceinline__ float to_elem<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 to_elem<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float ld_elem(const void* base, size_t idx, int dtype) {
  return dtype == DTG_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[idx])
                           : reinterpret_cast<const float*>(base)[idx];
}
__device__ __forceinline__ void st_elem(void* base, size_t idx, int dtype, float v) {
  if (dtype == DTG_BF16)
    reinterpret_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
  else
    reinterpret_cast<float*>(base)[idx] = round_tf32(v);
}

// one block-row (blockIdx.y) per item; threads stride over the padded destination
__global__ void pack_weights_kernel(const dtg_pack_item* items) {
  pdl_enter();
  const dtg_pack_item it = items[blockIdx.y];
  const int total = it.taps * it.rows_p * it.cols_p;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % it.cols_p;
    const int r = (i / it.cols_p) % it.rows_p;
    const int t = i / (it.cols_p * it.rows_p);
    float v = 0.f;
    if (it.s2d_k > 0) {
      // space-to-depth view of a stride-2, pad-1 KxK filter (K = 3 or 4) as a stride-1 3x3 filter over 2x2 pixel
      // blocks: tap (DY, DX), column (dy*2 + dx) * cp + b  <-  w[r][b][kh = 2*DY + dy - 1][kw = 2*DX + dx - 1]
      const int cp = it.fold_fc, K = it.s2d_k;
      const int sub = c / cp, b = c % cp;
      const int kh = 2 * (t / 3) + (sub >> 1) - 1, kw = 2 * (t % 3) + (sub & 1) - 1;
      if (r < it.rows && b < it.cols && sub < 4 && kh >= 0 && kh < K && kw >= 0 && kw < K)
        v = it.src[((static_cast<size_t>(r) * it.srs + static_cast<size_t>(b) * it.scs) * K + kh) * K + kw];
    } else if (it.fold_kw > 0 && it.fold_flip == 2) {
      // filter column folded into the ROWS (GEMM-N of conv_tail7.cu): row = kw * rows + r0, taps = KH
      const int j = r / it.rows, r0 = r % it.rows;
      if (j < it.fold_kw && c < it.cols)
        v = it.src[((static_cast<size_t>(r0) * it.srs + static_cast<size_t>(c) * it.scs) * it.taps + t) * it.fold_kw + j];
    } else if (it.fold_kw > 0) {
      const int j = c / it.fold_fc, b = c % it.fold_fc;
      if (r < it.rows && j < it.fold_kw && b < it.cols) {
        const int kw = it.fold_flip ? it.fold_kw - 1 - j : j;
        v = it.src[((static_cast<size_t>(r) * it.srs + static_cast<size_t>(b) * it.scs) * it.taps + t) * it.fold_kw + kw];
      }
    } else if (r < it.rows && c < it.cols) {
      v = it.src[(static_cast<size_t>(r) * it.srs + static_cast<size_t>(c) * it.scs) * it.taps + t];
    }
    st_elem(it.dst, i, it.dtype, v);
  }
}

// NCHW fp32 -> plane channels [c_off, c_off+c), mirrored into the halo.  One thread per (n,h,w).
__global__ void pack_nchw_kernel(const float* __restrict__ src, const float* __restrict__ tanh_y, int n, int c, int h, int w,
                                 dtg_plane dst, int c_off, int reflect) {
  pdl_enter();
  const size_t total = static_cast<size_t>(n) * h * w;
  const int Hb = dst.h + 2 * dst.halo, Wb = dst.w + 2 * dst.halo;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int x = i % w;
    const int y = (i / w) % h;
    const int b = i / (static_cast<size_t>(w) * h);
    int hts[3], wts[3];
    const int nh = reflect_targets(y, h, reflect ? dst.halo : 0, hts), nw = reflect_targets(x, w, reflect ? dst.halo : 0, wts);
    for (int ch = 0; ch < c; ++ch) {
      const size_t si = ((static_cast<size_t>(b) * c + ch) * h + y) * w + x;
      float v = src[si];
      if (tanh_y) {
        const float t = tanh_y[si];
        v *= (1.f - t * t);
      }
      for (int a = 0; a < nh; ++a)
        for (int q = 0; q < nw; ++q) {
          const size_t pix = (static_cast<size_t>(b) * Hb + hts[a] + dst.halo) * Wb + wts[q] + dst.halo;
          st_elem(dst.ptr, pix * dst.c + c_off + ch, dst.dtype, v);
        }
    }
  }
}

// NCHW fp32 -> space-to-depth plane [n][h/2][w/2][4*cp]: pixel (y, x) channel ch lands in block (y/2, x/2), channel
// ((y&1)*2 + (x&1)) * cp + c_off + ch.  One thread per (n, h, w).
__global__ void pack_nchw_s2d_kernel(const float* __restrict__ src, int n, int c, int h, int w, dtg_plane dst, int cp, int c_off) {
  pdl_enter();
  const size_t total = static_cast<size_t>(n) * h * w;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int x = i % w;
    const int y = (i / w) % h;
    const int b = i / (static_cast<size_t>(w) * h);
    const size_t pix = (static_cast<size_t>(b) * dst.h + (y >> 1)) * dst.w + (x >> 1);
    const int sub = (y & 1) * 2 + (x & 1);
    for (int ch = 0; ch < c; ++ch)
      st_elem(dst.ptr, pix * dst.c + sub * cp + c_off + ch, dst.dtype, src[((static_cast<size_t>(b) * c + ch) * h + y) * w + x]);
  }
}

// dw[co][b][kh][kw] += dw2[co][(dy*2+dx)*cp + b][DY][DX] for the space-to-depth filter view above; dw2 is cleared
// (every element, also the structurally-zero ones) so that the next accumulation starts from zero.
__global__ void s2d_unfold_add_kernel(float* __restrict__ dw2, float* __restrict__ dw, int cout, int cin, int K, int cp) {
  pdl_enter();
  const int total = cout * 4 * cp * 9;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int t = i % 9;
    const int col = (i / 9) % (4 * cp);
    const int co = i / (9 * 4 * cp);
    const int sub = col / cp, b = col % cp;
    const int kh = 2 * (t / 3) + (sub >> 1) - 1, kw = 2 * (t % 3) + (sub & 1) - 1;
    if (b < cin && kh >= 0 && kh < K && kw >= 0 && kw < K) dw[((static_cast<size_t>(co) * cin + b) * K + kh) * K + kw] += dw2[i];
    dw2[i] = 0.f;
  }
}

__global__ void unpack_nchw_kernel(dtg_plane src, int c_off, int c, float* __restrict__ dst) {
  pdl_enter();
  const size_t total = static_cast<size_t>(src.n) * src.h * src.w;
  const int Hb = src.h + 2 * src.halo, Wb = src.w + 2 * src.halo;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int x = i % src.w;
    const int y = (i / src.w) % src.h;
    const int b = i / (static_cast<size_t>(src.w) * src.h);
    const size_t pix = (static_cast<size_t>(b) * Hb + y + src.halo) * Wb + x + src.halo;
    for (int ch = 0; ch < c; ++ch)
      dst[((static_cast<size_t>(b) * c + ch) * src.h + y) * src.w + x] = ld_elem(src.ptr, pix * src.c + c_off + ch, src.dtype);
  }
}

// value of a gradient plane at interior pixel (y,x) channel ch with its halo folded back
__device__ __forceinline__ float folded_load(const dtg_plane& s, int b, int y, int x, int ch) {
  const int Hb = s.h + 2 * s.halo, Wb = s.w + 2 * s.halo;
  int hts[3], wts[3];
  const int nh = reflect_targets(y, s.h, s.halo, hts), nw = reflect_targets(x, s.w, s.halo, wts);
  float acc = 0.f;
  for (int a = 0; a < nh; ++a)
    for (int q = 0; q < nw; ++q) {
      const size_t pix = (static_cast<size_t>(b) * Hb + hts[a] + s.halo) * Wb + wts[q] + s.halo;
      acc += ld_elem(s.ptr, pix * s.c + ch, s.dtype);
    }
  return acc;
}

struct GatherArgs {
  dtg_plane src[3];
  int c_off[3];
  int nsrc;
};

__global__ void grad_gather_kernel(GatherArgs g, const float* __restrict__ add_nchw, const float* __restrict__ tanh_y, int c,
                                   dtg_plane out, float* __restrict__ out_nchw) {
  pdl_enter();
  const int h = out.h, w = out.w;
  const size_t total = static_cast<size_t>(out.n) * h * w;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int x = i % w;
    const int y = (i / w) % h;
    const int b = i / (static_cast<size_t>(w) * h);
    for (int ch = 0; ch < c; ++ch) {
      float acc = 0.f;
      for (int k = 0; k < g.nsrc; ++k) acc += folded_load(g.src[k], b, y, x, g.c_off[k] + ch);
      const size_t di = ((static_cast<size_t>(b) * c + ch) * h + y) * w + x;
      if (add_nchw) acc += add_nchw[di];
      if (out_nchw) out_nchw[di] = acc;
      if (tanh_y) {
        const float t = tanh_y[di];
        acc *= (1.f - t * t);
      }
      if (out.ptr) {
        const size_t pix = (static_cast<size_t>(b) * (h + 2 * out.halo) + y + out.halo) * (w + 2 * out.halo) + x + out.halo;
        st_elem(out.ptr, pix * out.c + ch, out.dtype, acc);
      }
    }
  }
}

// d_bias[c] += sum over (n,h,w) of a plane, c <= 16 (network heads).  Two-stage inside one launch: per-block
// partials, the last block to finish adds them up in a fixed order (deterministic, no float atomics).
constexpr int kCsBlocks = 96;
struct ChanSumWs {
  unsigned int counter;
  unsigned int pad[15];
  float part[kCsBlocks][16];
};

__global__ void __launch_bounds__(256) channel_sum_kernel(dtg_plane x, int c, float* __restrict__ d_bias, ChanSumWs* ws) {
  pdl_enter();
  const int Hb = x.h + 2 * x.halo, Wb = x.w + 2 * x.halo;
  const size_t total = static_cast<size_t>(x.n) * x.h * x.w;
  float acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.f;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int xx = i % x.w;
    const int yy = (i / x.w) % x.h;
    const int b = i / (static_cast<size_t>(x.w) * x.h);
    const size_t pix = (static_cast<size_t>(b) * Hb + yy + x.halo) * Wb + xx + x.halo;
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (k < c) acc[k] += ld_elem(x.ptr, pix * x.c + k, x.dtype);
  }
  __shared__ float red[8][16];
  __shared__ bool last;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    float v = acc[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
    ws->part[blockIdx.x][threadIdx.x] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int prev = atomicAdd(&ws->counter, 1u);
    last = prev == gridDim.x - 1;
    if (last) {
      ws->counter = 0;
      __threadfence();
    }
  }
  __syncthreads();
  if (last && threadIdx.x < c) {
    float v = 0.f;
    for (unsigned int b = 0; b < gridDim.x; ++b) v += ws->part[b][threadIdx.x];
    d_bias[threadIdx.x] += v;
  }
}

static inline int grid_for(size_t total, int block, int cap = 148 * 16) {
  size_t g = (total + block - 1) / block;
  if (g < 1) g = 1;
  if (g > static_cast<size_t>(cap)) g = cap;
  return static_cast<int>(g);
}

}  // namespace dtg

using namespace dtg;

extern "C" int dtg_pack_weights(const dtg_pack_item* items_dev, int nitems, int max_elems, void* stream) {
  DTG_REQUIRE(items_dev && nitems > 0, "dtg_pack_weights: no items");
  dim3 grid(grid_for(static_cast<size_t>(max_elems), 256, 64), nitems);
  DTG_CHECK_CUDA(launch_k(pack_weights_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), items_dev));
  return DTG_OK;
}

extern "C" int dtg_pack_nchw(const float* src, const float* tanh_y, int n, int c, int h, int w, const dtg_plane* dst, int c_off,
                             int reflect, void* stream) {
  DTG_REQUIRE(src && dst && dst->ptr, "dtg_pack_nchw: null");
  DTG_REQUIRE(dst->n == n && dst->h == h && dst->w == w && c_off + c <= dst->c, "dtg_pack_nchw: shape mismatch");
  const size_t total = static_cast<size_t>(n) * h * w;
  DTG_CHECK_CUDA(launch_k(pack_nchw_kernel, grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream), src, tanh_y, n, c, h, w, *dst, c_off, reflect));
  return DTG_OK;
}

extern "C" int dtg_pack_nchw_s2d(const float* src, int n, int c, int h, int w, const dtg_plane* dst, int cp, int c_off,
                                 void* stream) {
  DTG_REQUIRE(src && dst && dst->ptr, "dtg_pack_nchw_s2d: null");
  DTG_REQUIRE(h % 2 == 0 && w % 2 == 0 && dst->n == n && dst->h == h / 2 && dst->w == w / 2 && dst->halo == 0 &&
                  dst->c == 4 * cp && c_off + c <= cp,
              "dtg_pack_nchw_s2d: shape mismatch");
  const size_t total = static_cast<size_t>(n) * h * w;
  DTG_CHECK_CUDA(launch_k(pack_nchw_s2d_kernel, grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream), src, n, c, h, w, *dst, cp, c_off));
  return DTG_OK;
}

extern "C" int dtg_s2d_unfold_add(float* dw2, float* dw, int cout, int cin, int k, int cp, void* stream) {
  DTG_REQUIRE(dw2 && dw && cout > 0 && cin > 0 && cin <= cp && (k == 3 || k == 4), "dtg_s2d_unfold_add: bad args");
  const size_t total = static_cast<size_t>(cout) * 4 * cp * 9;
  DTG_CHECK_CUDA(launch_k(s2d_unfold_add_kernel, grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream), dw2, dw, cout, cin, k, cp));
  return DTG_OK;
}

extern "C" int dtg_unpack_nchw(const dtg_plane* src, int c_off, int c, float* dst, void* stream) {
  DTG_REQUIRE(src && src->ptr && dst && c_off + c <= src->c, "dtg_unpack_nchw: bad args");
  const size_t total = static_cast<size_t>(src->n) * src->h * src->w;
  DTG_CHECK_CUDA(launch_k(unpack_nchw_kernel, grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream), *src, c_off, c, dst));
  return DTG_OK;
}

extern "C" int dtg_grad_gather(const dtg_plane* const* srcs, const int* c_off, int nsrc, const float* add_nchw,
                               const float* tanh_y, int c, const dtg_plane* out, float* out_nchw, void* stream) {
  DTG_REQUIRE(srcs && nsrc >= 1 && nsrc <= 3 && out, "dtg_grad_gather: bad args");
  GatherArgs g;
  memset(&g, 0, sizeof(g));
  g.nsrc = nsrc;
  for (int k = 0; k < nsrc; ++k) {
    DTG_REQUIRE(srcs[k] && srcs[k]->ptr, "dtg_grad_gather: null source");
    DTG_REQUIRE(srcs[k]->n == out->n && srcs[k]->h == out->h && srcs[k]->w == out->w && c_off[k] + c <= srcs[k]->c,
                "dtg_grad_gather: source %d shape mismatch", k);
    g.src[k] = *srcs[k];
    g.c_off[k] = c_off[k];
  }
  DTG_REQUIRE(out->ptr == nullptr || c <= out->c, "dtg_grad_gather: out channels");
  const size_t total = static_cast<size_t>(out->n) * out->h * out->w;
  DTG_CHECK_CUDA(launch_k(grad_gather_kernel, grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream), g, add_nchw, tanh_y, c, *out, out_nchw));
  return DTG_OK;
}

extern "C" int dtg_channel_sum(const dtg_plane* x, int c, float* d_bias, void* workspace, void* stream) {
  DTG_REQUIRE(x && x->ptr && d_bias && workspace && c <= x->c && c <= 16, "dtg_channel_sum: bad args (c <= 16)");
  const size_t total = static_cast<size_t>(x->n) * x->h * x->w;
  const int blocks = static_cast<int>(std::max<size_t>(1, std::min<size_t>((total + 1023) / 1024, kCsBlocks)));
  DTG_CHECK_CUDA(launch_k(channel_sum_kernel, blocks, 256, 0, static_cast<cudaStream_t>(stream), *x, c, d_bias, reinterpret_cast<ChanSumWs*>(workspace)));
  return DTG_OK;
}
