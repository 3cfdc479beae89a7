// Fused loss reductions with backward seeds, and clip_grad_norm + Adam over flat fp32 arenas.
// All reductions are two-stage inside one launch (per-block partials, last block finishes in a fixed
// order) -> deterministic, no host synchronisation, CUDA-graph capturable.
#include <algorithm>

#include "common.cuh"

namespace dtg {

constexpr int kRedThreads = 256;
constexpr int kRedMaxBlocks = 120;  // workspace: 64-byte header + 120 * 8 floats <= 4 KB

struct RedWs {
  unsigned int counter;
  unsigned int pad[15];
  float part[kRedMaxBlocks][8];
};

// block-wide sum of up to K values per thread; result valid in thread 0
template <int K>
__device__ __forceinline__ void block_sum(float (&v)[K], float* sm) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    float x = v[k];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) sm[w * K + k] = x;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      float x = 0.f;
      for (int i = 0; i < kRedThreads / 32; ++i) x += sm[i * K + k];
      v[k] = x;
    }
  }
  __syncthreads();
}

// returns true in thread 0 of the LAST block to finish (partials of all blocks are then visible)
__device__ __forceinline__ bool publish_and_check_last(RedWs* ws, const float* vals, int k) {
  __shared__ bool last;
  if (threadIdx.x == 0) {
    for (int i = 0; i < k; ++i) ws->part[blockIdx.x][i] = vals[i];
    __threadfence();
    const unsigned int prev = atomicAdd(&ws->counter, 1u);
    last = (prev == gridDim.x - 1);
    if (last) {
      ws->counter = 0;  // self-reset for the next call on this stream
      __threadfence();
    }
  }
  __syncthreads();
  return last && threadIdx.x == 0;
}

__device__ __forceinline__ void st_plane(const dtg_plane& p, size_t idx, float v) {
  if (p.dtype == DTG_BF16)
    reinterpret_cast<__nv_bfloat16*>(p.ptr)[idx] = __float2bfloat16_rn(v);
  else
    reinterpret_cast<float*>(p.ptr)[idx] = round_tf32(v);
}

__device__ __forceinline__ void lsgan_body(const float* __restrict__ pred, int n, int h, int w, float target, float gscale,
                                           float* __restrict__ scalars, int slot_loss, int slot_mean, const dtg_plane& dp,
                                           RedWs* ws) {
  __shared__ float sm[(kRedThreads / 32) * 2];
  const int count = n * h * w;
  float v[2] = {0.f, 0.f};
  const float k = gscale * 2.f / static_cast<float>(count);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
    const float p = pred[i];
    const float d = p - target;
    v[0] += d * d;
    v[1] += p;
    if (dp.ptr) {
      const int x = i % w, y = (i / w) % h, b = i / (w * h);
      const size_t pix = (static_cast<size_t>(b) * (dp.h + 2 * dp.halo) + y + dp.halo) * (dp.w + 2 * dp.halo) + x + dp.halo;
      st_plane(dp, pix * dp.c, k * d);
    }
  }
  block_sum<2>(v, sm);
  if (publish_and_check_last(ws, v, 2)) {
    float a = 0.f, b = 0.f;
    for (unsigned int i = 0; i < gridDim.x; ++i) {
      a += ws->part[i][0];
      b += ws->part[i][1];
    }
    if (slot_loss >= 0) scalars[slot_loss] = a / static_cast<float>(count);
    if (slot_mean >= 0) scalars[slot_mean] = b / static_cast<float>(count);
  }
}

__global__ void __launch_bounds__(kRedThreads) lsgan_kernel(const float* __restrict__ pred, int n, int h, int w, float target,
                                                            float gscale, float* __restrict__ scalars, int slot_loss,
                                                            int slot_mean, dtg_plane dp, RedWs* ws) {
  pdl_enter();
  lsgan_body(pred, n, h, w, target, gscale, scalars, slot_loss, slot_mean, dp, ws);
}

__device__ __forceinline__ void l1_body(const float* __restrict__ a, const float* __restrict__ b, int n, int c, int h, int w,
                                        float gscale, int tanh_bwd, float* __restrict__ scalars, int slot_loss, int slot_aux,
                                        const dtg_plane& da, RedWs* ws) {
  __shared__ float sm[(kRedThreads / 32) * 2];
  __shared__ float smm[(kRedThreads / 32) * 2];
  const int count = n * c * h * w;
  float v[2] = {0.f, 0.f};
  float mn = INFINITY, mx = -INFINITY;
  const float k = gscale / static_cast<float>(count);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
    const float av = a[i];
    const float d = av - b[i];
    v[0] += fabsf(d);
    v[1] += av * av;
    mn = fminf(mn, av);
    mx = fmaxf(mx, av);
    if (da.ptr) {
      const int x = i % w, y = (i / w) % h, ch = (i / (w * h)) % c, bi = i / (w * h * c);
      float g = d > 0.f ? k : (d < 0.f ? -k : 0.f);
      if (tanh_bwd) g *= (1.f - av * av);
      const size_t pix = (static_cast<size_t>(bi) * (da.h + 2 * da.halo) + y + da.halo) * (da.w + 2 * da.halo) + x + da.halo;
      st_plane(da, pix * da.c + ch, g);
    }
  }
  block_sum<2>(v, sm);
  // block min / max
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) {
    smm[(threadIdx.x >> 5) * 2] = mn;
    smm[(threadIdx.x >> 5) * 2 + 1] = mx;
  }
  __syncthreads();
  float vals[4] = {v[0], v[1], 0.f, 0.f};
  if (threadIdx.x == 0) {
    for (int i = 0; i < kRedThreads / 32; ++i) {
      mn = fminf(mn, smm[i * 2]);
      mx = fmaxf(mx, smm[i * 2 + 1]);
    }
    vals[2] = mn;
    vals[3] = mx;
  }
  if (publish_and_check_last(ws, vals, 4)) {
    float s0 = 0.f, s1 = 0.f, gmn = INFINITY, gmx = -INFINITY;
    for (unsigned int i = 0; i < gridDim.x; ++i) {
      s0 += ws->part[i][0];
      s1 += ws->part[i][1];
      gmn = fminf(gmn, ws->part[i][2]);
      gmx = fmaxf(gmx, ws->part[i][3]);
    }
    if (slot_loss >= 0) scalars[slot_loss] = s0 / static_cast<float>(count);
    if (slot_aux >= 0) {
      scalars[slot_aux] = 0.5f * s1 / static_cast<float>(n);
      scalars[slot_aux + 1] = gmn;
      scalars[slot_aux + 2] = gmx;
    }
  }
}

__global__ void __launch_bounds__(kRedThreads) l1_kernel(const float* __restrict__ a, const float* __restrict__ b, int n, int c,
                                                         int h, int w, float gscale, int tanh_bwd,
                                                         float* __restrict__ scalars, int slot_loss, int slot_aux,
                                                         dtg_plane da, RedWs* ws) {
  pdl_enter();
  l1_body(a, b, n, c, h, w, gscale, tanh_bwd, scalars, slot_loss, slot_aux, da, ws);
}

// Multi-segment loss reduction: blockIdx.y selects one loss term (LSGAN MSE against a constant target, or L1 with the
// optional tanh-backward factor and the KLD / min / max side outputs); every term reduces into its own slots of the
// packed scalar vector and writes its own seed-gradient plane, all in ONE launch (model.py:432-439, 458-505).
constexpr int kMaxLossSegs = 8;
struct LossSegDev {
  const float* a;
  const float* b;
  int kind, n, c, h, w;
  float target, gscale;
  int tanh_bwd, slot_loss, slot_aux;
  dtg_plane d;
};
struct LossSegs {
  LossSegDev s[kMaxLossSegs];
};

__global__ void __launch_bounds__(kRedThreads) loss_fused_kernel(const __grid_constant__ LossSegs L, float* __restrict__ scalars,
                                                                 RedWs* ws) {
  pdl_enter();
  const LossSegDev& g = L.s[blockIdx.y];
  if (g.kind == 0)
    lsgan_body(g.a, g.n, g.h, g.w, g.target, g.gscale, scalars, g.slot_loss, g.slot_aux, g.d, ws + blockIdx.y);
  else
    l1_body(g.a, g.b, g.n, g.c, g.h, g.w, g.gscale, g.tanh_bwd, scalars, g.slot_loss, g.slot_aux, g.d, ws + blockIdx.y);
}

__global__ void __launch_bounds__(kRedThreads) sumsq_kernel(const float* __restrict__ g, size_t count, float gscale,
                                                            float* __restrict__ out, RedWs* ws) {
  pdl_enter();
  __shared__ float sm[kRedThreads / 32];
  float v[1] = {0.f};
  const size_t n4 = count / 4;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 q = g4[i];
    v[0] += q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
  }
  if (blockIdx.x == 0)
    for (size_t i = n4 * 4 + threadIdx.x; i < count; i += blockDim.x) v[0] += g[i] * g[i];
  block_sum<1>(v, sm);
  if (publish_and_check_last(ws, v, 1)) {
    double a = 0.0;
    for (unsigned int i = 0; i < gridDim.x; ++i) a += ws->part[i][0];
    *out = static_cast<float>(a) * gscale * gscale;
  }
}

__global__ void __launch_bounds__(256) adam_clip_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, size_t count, const float* __restrict__ hyper,
                                                        const float* __restrict__ sumsq, const int32_t* __restrict__ step_dev,
                                                        float gscale) {
  pdl_enter();
  __shared__ float s_coef, s_step_size, s_inv_sqrt_bc2;
  if (threadIdx.x == 0) {
    const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], max_norm = hyper[4];
    const float norm = sqrtf(*sumsq);
    const float cc = max_norm / (norm + 1e-6f);
    s_coef = (cc < 1.f ? cc : 1.f) * gscale;
    const double t = static_cast<double>(*step_dev);
    const double bc1 = 1.0 - pow(static_cast<double>(b1), t);
    const double bc2 = 1.0 - pow(static_cast<double>(b2), t);
    s_step_size = static_cast<float>(static_cast<double>(lr) / bc1);
    s_inv_sqrt_bc2 = static_cast<float>(1.0 / sqrt(bc2));
  }
  __syncthreads();
  const float b1 = hyper[1], b2 = hyper[2], eps = hyper[3];
  const float coef = s_coef, step_size = s_step_size, isb = s_inv_sqrt_bc2;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < count; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float gi = g[i] * coef;
    g[i] = gi;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= step_size * mi / (sqrtf(vi) * isb + eps);
  }
}

__global__ void step_inc_kernel(int32_t* s) {
  pdl_enter(); *s += 1; }

// ---- evaluate.py:89-124 (variational_ubo): the objective around G_A_B and the RMSprop step on q(z) -------------------
// Laplace log-likelihood of real_B under (fake_B, logvar_B) (model.py:24-28), summed over everything and divided by n
// (= mean over samples of the per-sample sums, evaluate.py:93-94, 118), plus the seed gradient of
// loss = mean_n(-log_prob_n) with respect to the generator's pre-tanh output.
__global__ void __launch_bounds__(kRedThreads) ubo_laplace_kernel(const float* __restrict__ fake, const float* __restrict__ real,
                                                                  const float* __restrict__ logvar_b, int n, int c, int h, int w,
                                                                  float* __restrict__ scalars, int slot_logp, dtg_plane df,
                                                                  RedWs* ws) {
  pdl_enter();
  __shared__ float sm[kRedThreads / 32];
  const int chw = c * h * w, count = n * chw;
  const float k = 1.f / static_cast<float>(n);
  float v[1] = {0.f};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
    const float fv = fake[i];
    const float d = real[i] - fv;
    const float lv = logvar_b[i % chw];
    const float isd = __expf(-0.5f * lv);
    v[0] += -0.5f * lv - fabsf(d) * isd - 0.69314718056f;
    if (df.ptr) {
      const int x = i % w, y = (i / w) % h, ch = (i / (w * h)) % c, bi = i / chw;
      // d(-log_prob)/d fake = -sign(real - fake) / sd, through tanh
      float g = d > 0.f ? -k * isd : (d < 0.f ? k * isd : 0.f);
      g *= (1.f - fv * fv);
      const size_t pix = (static_cast<size_t>(bi) * (df.h + 2 * df.halo) + y + df.halo) * (df.w + 2 * df.halo) + x + df.halo;
      st_plane(df, pix * df.c + ch, g);
    }
  }
  block_sum<1>(v, sm);
  if (publish_and_check_last(ws, v, 1)) {
    double a = 0.0;
    for (unsigned int i = 0; i < gridDim.x; ++i) a += ws->part[i][0];
    scalars[slot_logp] = static_cast<float>(a / n);
  }
}

// One block: KL(q || N(0,I)) of the CURRENT (mu, logvar) (model.py:45-53, mean over samples), the gradients of
// loss = mean_n(-log_prob_n + kld_n) with respect to (mu, logvar) from dz (through z = clamp(mu + eps * sd, -4, 4),
// model.py:15-22), torch.optim.RMSprop's update (alpha, eps; no momentum, not centered; evaluate.py:65, 119-121) and the
// next iterate z' = clamp(mu' + eps_next * sd', -4, 4) (evaluate.py:123).
__global__ void __launch_bounds__(kRedThreads) ubo_latent_step_kernel(float* __restrict__ mu, float* __restrict__ logvar,
                                                                      float* __restrict__ sq_mu, float* __restrict__ sq_lv,
                                                                      const float* __restrict__ eps_cur,
                                                                      const float* __restrict__ eps_next,
                                                                      const float* __restrict__ dz, int n, int nz, float lr,
                                                                      float alpha, float rms_eps, float* __restrict__ z_out,
                                                                      float* __restrict__ scalars, int slot_kld) {
  pdl_enter();
  __shared__ float sm[kRedThreads / 32];
  const int count = n * nz;
  const float inv_n = 1.f / static_cast<float>(n);
  float v[1] = {0.f};
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    float m = mu[i], lv = logvar[i];
    const float ex = expf(lv), sd = expf(0.5f * lv);
    v[0] += -0.5f * (lv + 1.f - m * m - ex);
    const float e = eps_cur[i];
    const float zr = m + e * sd;
    const float g = (zr >= -4.f && zr <= 4.f) ? dz[i] : 0.f;
    const float g_mu = g + m * inv_n;
    const float g_lv = g * e * 0.5f * sd - 0.5f * (1.f - ex) * inv_n;
    const float s1 = alpha * sq_mu[i] + (1.f - alpha) * g_mu * g_mu;
    const float s2 = alpha * sq_lv[i] + (1.f - alpha) * g_lv * g_lv;
    sq_mu[i] = s1;
    sq_lv[i] = s2;
    m -= lr * g_mu / (sqrtf(s1) + rms_eps);
    lv -= lr * g_lv / (sqrtf(s2) + rms_eps);
    mu[i] = m;
    logvar[i] = lv;
    z_out[i] = fminf(4.f, fmaxf(-4.f, m + eps_next[i] * expf(0.5f * lv)));
  }
  block_sum<1>(v, sm);
  if (threadIdx.x == 0) scalars[slot_kld] = v[0] * inv_n;
}

}  // namespace dtg

using namespace dtg;

static int red_blocks(size_t count) {
  return static_cast<int>(std::max<size_t>(1, std::min<size_t>((count + kRedThreads * 4 - 1) / (kRedThreads * 4), kRedMaxBlocks)));
}

extern "C" int dtg_loss_lsgan(const float* pred, int n, int h, int w, float target, float grad_scale, float* scalars,
                              int slot_loss, int slot_mean, const dtg_plane* dpred, void* workspace, void* stream) {
  DTG_REQUIRE(pred && scalars && workspace, "dtg_loss_lsgan: null argument");
  dtg_plane dp = {nullptr, 0, 0, 0, 0, 0, 0};
  if (dpred && dpred->ptr) {
    DTG_REQUIRE(dpred->n == n && dpred->h == h && dpred->w == w, "dtg_loss_lsgan: dpred plane mismatch");
    dp = *dpred;
  }
  DTG_CHECK_CUDA(launch_k(lsgan_kernel, red_blocks(static_cast<size_t>(n) * h * w), kRedThreads, 0, static_cast<cudaStream_t>(stream), pred, n, h, w, target, grad_scale, scalars, slot_loss, slot_mean, dp, reinterpret_cast<RedWs*>(workspace)));
  return DTG_OK;
}

extern "C" int dtg_loss_l1(const float* a, const float* b, int n, int c, int h, int w, float grad_scale, int tanh_bwd,
                           float* scalars, int slot_loss, int slot_aux, const dtg_plane* da, void* workspace, void* stream) {
  DTG_REQUIRE(a && b && scalars && workspace, "dtg_loss_l1: null argument");
  dtg_plane dp = {nullptr, 0, 0, 0, 0, 0, 0};
  if (da && da->ptr) {
    DTG_REQUIRE(da->n == n && da->h == h && da->w == w && da->c >= c, "dtg_loss_l1: da plane mismatch");
    dp = *da;
  }
  DTG_CHECK_CUDA(launch_k(l1_kernel, red_blocks(static_cast<size_t>(n) * c * h * w), kRedThreads, 0, static_cast<cudaStream_t>(stream), a, b, n, c, h, w, grad_scale, tanh_bwd, scalars, slot_loss, slot_aux, dp, reinterpret_cast<RedWs*>(workspace)));
  return DTG_OK;
}

extern "C" int dtg_loss_fused(const dtg_loss_seg* segs, int nseg, float* scalars, void* workspace, void* stream) {
  DTG_REQUIRE(segs && scalars && workspace && nseg >= 1 && nseg <= kMaxLossSegs, "dtg_loss_fused: 1..%d segments", kMaxLossSegs);
  LossSegs L;
  memset(&L, 0, sizeof(L));
  size_t most = 1;
  for (int i = 0; i < nseg; ++i) {
    const dtg_loss_seg& q = segs[i];
    DTG_REQUIRE(q.a && (q.kind == DTG_LOSS_LSGAN || (q.kind == DTG_LOSS_L1 && q.b)), "dtg_loss_fused: segment %d: null tensor / bad kind", i);
    LossSegDev& d = L.s[i];
    d.a = q.a;
    d.b = q.b;
    d.kind = q.kind;
    d.n = q.n;
    d.c = q.kind == DTG_LOSS_LSGAN ? 1 : q.c;
    d.h = q.h;
    d.w = q.w;
    d.target = q.target;
    d.gscale = q.grad_scale;
    d.tanh_bwd = q.tanh_bwd;
    d.slot_loss = q.slot_loss;
    d.slot_aux = q.slot_aux;
    if (q.grad && q.grad->ptr) {
      DTG_REQUIRE(q.grad->n == q.n && q.grad->h == q.h && q.grad->w == q.w && q.grad->c >= d.c, "dtg_loss_fused: segment %d: gradient plane mismatch", i);
      d.d = *q.grad;
    }
    most = std::max(most, static_cast<size_t>(d.n) * d.c * d.h * d.w);
  }
  DTG_CHECK_CUDA(launch_k(loss_fused_kernel, dim3(red_blocks(most), nseg, 1), kRedThreads, 0, static_cast<cudaStream_t>(stream), L, scalars,
                          reinterpret_cast<RedWs*>(workspace)));
  return DTG_OK;
}

extern "C" int dtg_grad_sumsq(const float* g, size_t count, float grad_scale, float* out_sumsq, void* workspace, void* stream) {
  DTG_REQUIRE(g && out_sumsq && workspace, "dtg_grad_sumsq: null argument");
  DTG_REQUIRE((reinterpret_cast<uintptr_t>(g) & 15) == 0, "dtg_grad_sumsq: arena must be 16-byte aligned");
  DTG_CHECK_CUDA(launch_k(sumsq_kernel, red_blocks(count / 4), kRedThreads, 0, static_cast<cudaStream_t>(stream), g, count, grad_scale, out_sumsq,
                                                                                              reinterpret_cast<RedWs*>(workspace)));
  return DTG_OK;
}

extern "C" int dtg_adam_clip(float* p, float* g, float* m, float* v, size_t count, const float* hyper, const float* sumsq,
                             const int32_t* step_dev, float grad_scale, void* stream) {
  DTG_REQUIRE(p && g && m && v && hyper && sumsq && step_dev, "dtg_adam_clip: null argument");
  const int grid = static_cast<int>(std::max<size_t>(1, std::min<size_t>((count + 255) / 256, 148 * 8)));
  DTG_CHECK_CUDA(launch_k(adam_clip_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), p, g, m, v, count, hyper, sumsq, step_dev, grad_scale));
  return DTG_OK;
}

extern "C" int dtg_step_increment(int32_t* step_dev, void* stream) {
  DTG_REQUIRE(step_dev, "dtg_step_increment: null");
  DTG_CHECK_CUDA(launch_k(step_inc_kernel, 1, 1, 0, static_cast<cudaStream_t>(stream), step_dev));
  return DTG_OK;
}

extern "C" int dtg_ubo_laplace(const float* fake, const float* real, const float* logvar_b, int n, int c, int h, int w,
                               float* scalars, int slot_logp, const dtg_plane* dfake, void* workspace, void* stream) {
  DTG_REQUIRE(fake && real && logvar_b && scalars && workspace && slot_logp >= 0, "dtg_ubo_laplace: null argument");
  dtg_plane d;
  memset(&d, 0, sizeof(d));
  if (dfake && dfake->ptr) {
    DTG_REQUIRE(dfake->n == n && dfake->h == h && dfake->w == w && dfake->c >= c, "dtg_ubo_laplace: gradient plane mismatch");
    d = *dfake;
  }
  DTG_CHECK_CUDA(launch_k(ubo_laplace_kernel, red_blocks(static_cast<size_t>(n) * c * h * w), kRedThreads, 0,
                          static_cast<cudaStream_t>(stream), fake, real, logvar_b, n, c, h, w, scalars, slot_logp, d,
                          reinterpret_cast<RedWs*>(workspace)));
  return DTG_OK;
}

extern "C" int dtg_ubo_latent_step(float* mu, float* logvar, float* sq_mu, float* sq_logvar, const float* eps_cur,
                                   const float* eps_next, const float* dz, int n, int nz, float lr, float alpha,
                                   float rms_eps, float* z_out, float* scalars, int slot_kld, void* stream) {
  DTG_REQUIRE(mu && logvar && sq_mu && sq_logvar && eps_cur && eps_next && dz && z_out && scalars && slot_kld >= 0 && n > 0 && nz > 0,
              "dtg_ubo_latent_step: null argument");
  DTG_CHECK_CUDA(launch_k(ubo_latent_step_kernel, 1, kRedThreads, 0, static_cast<cudaStream_t>(stream), mu, logvar, sq_mu, sq_logvar,
                          eps_cur, eps_next, dz, n, nz, lr, alpha, rms_eps, z_out, scalars, slot_kld));
  return DTG_OK;
}
