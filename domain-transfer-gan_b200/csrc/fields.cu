// .npz field preprocessing of the reference's loader (dataloader.py:17-34) on the device: first `c` channels of an
// NHWC stack, NaN -> 0 (np.nan_to_num), per-sample / per-channel min-max scaling to [-1, 1] (constant fields -> 0),
// optional resize to gh x gw with skimage.transform.resize's defaults of the releases that still ran on Python 2
// (<= 0.14: order 1, mode 'constant' cval 0, no anti-aliasing; input coordinate = scale * (o + 0.5) - 0.5, samples
// outside the image contribute 0), NHWC -> NCHW float32.  HBM-bound: one read of the stack for the extrema, one
// gather for the resample; the scaling arithmetic runs in the array's own dtype like numpy's, the interpolation in
// double like skimage's.
#include <float.h>

#include <algorithm>

#include "common.cuh"

namespace dtg {

template <typename T>
struct FieldLim;
template <>
struct FieldLim<float> {
  static __device__ __forceinline__ float big() { return FLT_MAX; }
};
template <>
struct FieldLim<double> {
  static __device__ __forceinline__ double big() { return DBL_MAX; }
};

// np.nan_to_num: NaN -> 0, +-inf -> +-largest finite
template <typename T>
__device__ __forceinline__ T nan_to_num(T v) {
  if (v != v) return T(0);
  if (v > FieldLim<T>::big()) return FieldLim<T>::big();
  if (v < -FieldLim<T>::big()) return -FieldLim<T>::big();
  return v;
}

// grid (c, b): extrema of channel ch of sample b over h x w -> lohi[(b * c + ch) * 2 + {0, 1}]
template <typename T>
__global__ void __launch_bounds__(256) fields_minmax_kernel(const T* __restrict__ src, int h, int w, int cin, int c,
                                                            T* __restrict__ lohi) {
  pdl_enter();
  __shared__ T s_lo[8], s_hi[8];
  const int ch = blockIdx.x, b = blockIdx.y;
  const T* p = src + static_cast<size_t>(b) * h * w * cin + ch;
  T lo = FieldLim<T>::big(), hi = -FieldLim<T>::big();
  for (int i = threadIdx.x; i < h * w; i += blockDim.x) {
    const T v = nan_to_num(p[static_cast<size_t>(i) * cin]);
    lo = v < lo ? v : lo;
    hi = v > hi ? v : hi;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const T l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
    lo = l2 < lo ? l2 : lo;
    hi = h2 > hi ? h2 : hi;
  }
  if ((threadIdx.x & 31) == 0) {
    s_lo[threadIdx.x >> 5] = lo;
    s_hi[threadIdx.x >> 5] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) {
      lo = s_lo[i] < lo ? s_lo[i] : lo;
      hi = s_hi[i] > hi ? s_hi[i] : hi;
    }
    lohi[(static_cast<size_t>(b) * c + ch) * 2] = lo;
    lohi[(static_cast<size_t>(b) * c + ch) * 2 + 1] = hi;
  }
}

// the loader's scaling in the array's dtype and numpy's evaluation order: -1 + 2 * (v - lo) / (hi - lo); 0/0, x/0 -> 0
template <typename T>
__device__ __forceinline__ double scaled(const T* __restrict__ p, int y, int x, int h, int w, int cin, T lo, T hi) {
  if (y < 0 || y >= h || x < 0 || x >= w) return 0.0;      // mode 'constant', cval 0
  const T v = nan_to_num(p[(static_cast<size_t>(y) * w + x) * cin]);
  const T d = hi - lo;
  if (d == T(0)) return 0.0;
  const T s = T(-1) + T(2) * (v - lo) / d;
  if (s != s || s > FieldLim<T>::big() || s < -FieldLim<T>::big()) return 0.0;
  return static_cast<double>(s);
}

template <typename T>
__global__ void __launch_bounds__(256) fields_resample_kernel(const T* __restrict__ src, const T* __restrict__ lohi, int b,
                                                              int h, int w, int cin, int c, int gh, int gw,
                                                              float* __restrict__ dst) {
  pdl_enter();
  const size_t total = static_cast<size_t>(b) * c * gh * gw;
  const bool resize = gh != h || gw != w;
  const double rs = static_cast<double>(h) / gh, cs = static_cast<double>(w) / gw;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int ox = static_cast<int>(i % gw), oy = static_cast<int>((i / gw) % gh);
    const int ch = static_cast<int>((i / (static_cast<size_t>(gw) * gh)) % c), bi = static_cast<int>(i / (static_cast<size_t>(gw) * gh * c));
    const T* p = src + static_cast<size_t>(bi) * h * w * cin + ch;
    const T lo = lohi[(static_cast<size_t>(bi) * c + ch) * 2], hi = lohi[(static_cast<size_t>(bi) * c + ch) * 2 + 1];
    double out;
    if (!resize) {
      out = scaled(p, oy, ox, h, w, cin, lo, hi);
    } else {
      const double r = rs * (oy + 0.5) - 0.5, q = cs * (ox + 0.5) - 0.5;
      const double fr = floor(r), fq = floor(q);
      const int r0 = static_cast<int>(fr), q0 = static_cast<int>(fq);
      const int r1 = static_cast<int>(ceil(r)), q1 = static_cast<int>(ceil(q));
      const double dr = r - fr, dq = q - fq;
      const double top = (1.0 - dq) * scaled(p, r0, q0, h, w, cin, lo, hi) + dq * scaled(p, r0, q1, h, w, cin, lo, hi);
      const double bot = (1.0 - dq) * scaled(p, r1, q0, h, w, cin, lo, hi) + dq * scaled(p, r1, q1, h, w, cin, lo, hi);
      out = (1.0 - dr) * top + dr * bot;
    }
    dst[i] = static_cast<float>(out);
  }
}

template <typename T>
static int run_fields(const T* src, int b, int h, int w, int cin, int c, int gh, int gw, float* dst, T* lohi, cudaStream_t stream) {
  DTG_CHECK_CUDA(launch_k(fields_minmax_kernel<T>, dim3(c, b, 1), 256, 0, stream, src, h, w, cin, c, lohi));
  const size_t total = static_cast<size_t>(b) * c * gh * gw;
  const int grid = static_cast<int>(std::max<size_t>(1, std::min<size_t>((total + 255) / 256, 148 * 16)));
  DTG_CHECK_CUDA(launch_k(fields_resample_kernel<T>, grid, 256, 0, stream, src, static_cast<const T*>(lohi), b, h, w, cin, c, gh, gw, dst));
  return DTG_OK;
}

}  // namespace dtg

using namespace dtg;

extern "C" int dtg_preprocess_fields(const void* src, int is_f64, int b, int h, int w, int cin, int c, int gh, int gw,
                                     float* dst, void* lohi, void* stream) {
  DTG_REQUIRE(src && dst && lohi && b > 0 && h > 0 && w > 0 && c > 0 && c <= cin && gh > 0 && gw > 0 && b <= 65535,
              "dtg_preprocess_fields: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (is_f64) return run_fields(static_cast<const double*>(src), b, h, w, cin, c, gh, gw, dst, static_cast<double*>(lohi), s);
  return run_fields(static_cast<const float*>(src), b, h, w, cin, c, gh, gw, dst, static_cast<float*>(lohi), s);
}
