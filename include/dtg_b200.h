/* dtg_b200.h -- C ABI of the B200-native (sm_100a) Augmented-CycleGAN hot-path kernels.
 *
 * The reference (adrianalbert/domain-transfer-GAN) has NO FFI / plugin interface: its seam is the
 * Python nn.Module API and every arithmetic op is a PyTorch library call.  Each entry point below
 * names the reference call site(s) whose library kernel it replaces (paths relative to
 * /root/reference/augmented_cyclegan/).  The Python modules in domain-transfer-gan_b200/ bind these
 * with ctypes (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *  - extern "C", plain structs, raw device pointers, no C++ types; every function returns int:
 *    0 = DTG_OK, <0 = error; message via dtg_last_error().  Never aborts, never throws.
 *  - The caller owns ALL device memory (activations, workspaces, statistics).  The library keeps no
 *    device state; work is enqueued on the caller's stream without synchronisation, so every call
 *    is CUDA-graph capturable.  `stream` is a cudaStream_t passed as void*.
 *  - Activations are "planes": NHWC with an optional halo ring, element type bf16 or fp32
 *    (fp32 planes feed tcgen05 kind::tf32).  Buffer shape [n][h+2*halo][w+2*halo][c]; c*elsize must
 *    be a multiple of 16 bytes and ptr 128-byte aligned.  A halo > 0 holds MATERIALISED padding
 *    (reflection padding written by the producer); zero padding needs no halo (TMA out-of-bounds
 *    fill).
 */
#ifndef DTG_B200_H_
#define DTG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DTG_VERSION 124 /* 124: dtg_conv fold_w = 2; 123: dtg_preprocess_fields; 122: dtg_ubo_laplace, dtg_ubo_latent_step; 121: dtg_norm_bwd phases 3 / 4; 120: dtg_set_option, dtg_loss_fused; two-phase / TMA-staged instance-norm kernels behind dtg_norm_fwd / dtg_norm_bwd */

enum { DTG_OK = 0, DTG_ERR_INVALID = -1, DTG_ERR_CUDA = -2, DTG_ERR_UNSUPPORTED = -3 };
enum { DTG_BF16 = 0, DTG_F32 = 1 };
enum { DTG_ACT_NONE = 0, DTG_ACT_RELU = 1, DTG_ACT_LRELU = 2, DTG_ACT_TANH = 3 };
/* DTG_NORM_INSTANCE: biased variance (modules.py:83-97); DTG_NORM_COND_INSTANCE: unbiased variance,
 * per-sample affine (modules.py:120-132); DTG_NORM_BATCH: nn.BatchNorm{1,2}d training mode
 * (networks.py:407-415,450-462); DTG_NORM_NONE: activation only. */
enum { DTG_NORM_NONE = 0, DTG_NORM_INSTANCE = 1, DTG_NORM_COND_INSTANCE = 2, DTG_NORM_BATCH = 3 };
enum { DTG_CONV_FWD = 0, DTG_CONV_DGRAD = 1 };

typedef struct dtg_plane {
  void* ptr;
  int32_t n, h, w, c; /* interior extents; c = stored (padded) channels */
  int32_t halo;
  int32_t dtype; /* DTG_BF16 / DTG_F32 */
} dtg_plane;

int dtg_version(void);
/* number of kernels this library has enqueued so far in this process (monotonic; for bench accounting) */
unsigned long long dtg_launch_count(void);
/* copies the calling thread's last error message (NUL terminated) into buf; returns its length */
int dtg_last_error(char* buf, size_t cap);
/* Runtime options (process-wide; each also has an environment default read at first use).  Returns the previous value,
 * or DTG_ERR_INVALID for an unknown key.
 *   "pdl"         1 (default; env DTG_NO_PDL=1 -> 0): kernels are launched with the programmatic-dependent-launch
 *                 attribute and start with griddepcontrol.launch_dependents / .wait, so a prologue overlaps the previous
 *                 kernel's tail.  0 = plain stream-ordered launches: per-kernel device durations (CUPTI, ncu) then do not
 *                 include time spent waiting for a predecessor (bench.py's per-kernel roofline pass).
 *   "norm_impl"   which kernels serve the instance / conditional-instance modes of dtg_norm_fwd / dtg_norm_bwd where the
 *                 geometry allows (env DTG_NORM_IMPL): 2 (default) two-phase streaming forward (norm_lean.cu) + register-
 *                 resident cluster backward (norm_fused.cu) -- the fastest pair measured on B200; 3 two-phase streaming
 *                 forward and backward; 1 TMA-staged cluster kernels (norm_tma.cu); 0 cluster kernels of norm_fused.cu.
 *   "smem_cap_kb" 227 (env DTG_SMEM_CAP_KB): shared-memory budget of one tensor-core CTA; what it leaves free decides
 *                 whether bandwidth-bound kernels of other streams can be co-resident on the same SM.
 * (A non-deterministic red.global.add mode for the weight-gradient split-K partials, which SURVEY 7.1 allows, was NOT
 * added: the deterministic reduction -- fixed-order tree, one red.add owner per address -- costs 6.6 us per layer.) */
int dtg_set_option(const char* key, int value);

/* ---------------------------------------------------------------------------------------------
 * Weight packing.  Replaces cuDNN's internal filter transforms for nn.Conv2d / nn.ConvTranspose2d /
 * nn.Linear weights (networks.py:159-188,211-243,322-337,366-381,405-419,446-471).
 * dst[t][r][c] (rows padded to rows_p with zeros, cols to cols_p) = src[(r*srs + c*scs)*taps + t]
 * for r < rows, c < cols.  One item per launch-y; items live in DEVICE memory.
 * kw-folded form (fold_kw = KW > 0; small-channel 7x7 layers, see dtg_conv fold_w): `taps` = KH and a
 * destination column c = j*fold_fc + b holds filter column j (KW-1-j if fold_flip) of inner channel b:
 * dst[kh][r][j*fold_fc + b] = src[((r*srs + b*scs)*KH + kh)*KW + kw(j)] for j < KW, b < cols, else 0.
 * fold_flip = 2: the filter column goes into the ROWS instead (dtg_conv fold_w = 2): `taps` = KH,
 * dst[kh][kw*rows + r][c] = src[((r*srs + c*scs)*KH + kh)*KW + kw], rows_p >= KW*rows.
 * ------------------------------------------------------------------------------------------- */
typedef struct dtg_pack_item {
  const float* src;
  void* dst;
  int32_t rows, rows_p, cols, cols_p, taps, srs, scs, dtype;
  int32_t fold_kw, fold_flip, fold_fc;
  /* K > 0: space-to-depth packing of a stride-2, pad-1 KxK filter (K = 3 or 4) whose input has <= fold_fc channel
   * slots per pixel (fold_fc = channels of the 16..64-byte pixel): taps = 9, dst[DY*3+DX][r][(dy*2+dx)*fold_fc + b] =
   * w[r][b][2*DY+dy-1][2*DX+dx-1] (zero outside the filter).  With the input stored by dtg_pack_nchw_s2d the
   * strided convolution becomes a stride-1 3x3 convolution over 2x2 pixel blocks (first layers of
   * Discriminator / Discriminator_edges / LatentEncoder, networks.py:322,366,446). */
  int32_t s2d_k;
} dtg_pack_item;
int dtg_pack_weights(const dtg_pack_item* items_dev, int nitems, int max_elems, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution on tcgen05 tensor cores (TMA-staged NHWC tiles, TMEM accumulators).
 *   DTG_CONV_FWD  : y = act(conv(x, w) + bias)            nn.Conv2d forward
 *                   (modules.py:162,180,211,227; networks.py:160-187,212-242,322-337,366-381,446-471)
 *                   and the data gradient of nn.ConvTranspose2d.
 *   DTG_CONV_DGRAD: dx = conv_transpose(dy, w)            cuDNN dgrad of the above, and the FORWARD
 *                   of nn.ConvTranspose2d(k3,s2,p1,op1) (networks.py:178-179,231-234); stride-2 is
 *                   computed as 4 output-parity sub-convolutions (no zero insertion).
 * `w` is packed by dtg_pack_weights as [kh*kw][w_rows][w_cols]: rows = GEMM-N (output channels of
 * THIS call, multiple of 16), cols = GEMM-K per tap (input channels of this call).  The tap index
 * is always kh*KW+kw of the underlying correlation (no flipping in memory).
 * in.halo > 0 means the padding is materialised in the halo (must be >= pad for FWD).
 * DGRAD `ring`: also compute `ring` halo rings of the output (gradient w.r.t. a reflect-padded
 * input; the consumer folds them back); out.halo must be >= ring.  `in` (dy) has halo 0, or -- stride 1, ring > 0,
 * full 128-byte channel chunks, <= 128 output channels, out.halo == ring, n*(h+2r)*(w+2r) divisible by 8, w + 2r <= 63 -- a
 * ZERO halo ring == ring: the flat-raster path (conv_patch2.cu) then tiles all images as one tall image.
 * Output: a plane (dtype of `in`), optionally mirrored into its halo (out_reflect, reflection
 * padding for the next conv), or a dense fp32 NCHW tensor [n][cout][oh][ow] (out_nchw_f32 = 1,
 * cout <= 16).
 * fold_w = 1 (stride 1, ring 0): `in` is a 16-byte-per-pixel plane (8 bf16 / 4 fp32 channels) with a
 * MATERIALISED halo >= pad; the kernel reads it through an overlapping-row TMA view (row of pixel p =
 * the 128 bytes of pixels p..p+7), i.e. the KW filter columns become GEMM-K: taps = KH only and
 * `w` is the kw-folded packing [kh][w_rows][128 bytes] (dtg_pack_item fold_kw; fold_flip for DGRAD).
 * This is the generators' 7x7 head (networks.py:159-160,211-212) and the data gradient of their 7x7
 * tail (networks.py:187,242) without a materialised im2col.
 * fold_w = 2 (FWD, stride 1, "same" odd kernel, cout <= 4 with kw*cout <= 28, input halo 0 and 32 / 64 / 128 bytes
 * per pixel, width a divisor of 128, out_nchw_f32): the filter COLUMN goes into GEMM-N -- `w` is the
 * [kh][32][cin] packing of fold_flip = 2, only the KH filter rows are taps (14 tcgen05.mma per 128-pixel tile
 * instead of 98 for 32 -> 3 channels) and the epilogue adds the KW column partials with their pixel shift.
 * This is the forward of the generators' 7x7 tail + tanh (networks.py:187-188, 242-243).  With mode DGRAD (ring == pad,
 * `out` a 16-byte-pixel plane with halo == ring, `w` = [kh][(kw, cin) padded to 32][cout]) the same kernel computes the
 * data gradient of the generators' 7x7 HEAD w.r.t. its reflect-padded input (networks.py:159-160, 211-212): the "full"
 * correlation, (h + kh - 1) x (w + kw - 1) outputs per image, taps and column shifts mirrored.
 * ------------------------------------------------------------------------------------------- */
typedef struct dtg_conv_args {
  int32_t mode; /* DTG_CONV_FWD / DTG_CONV_DGRAD */
  int32_t kh, kw, stride, pad;
  int32_t ring;         /* DGRAD only */
  int32_t act;          /* DTG_ACT_* applied after bias */
  int32_t cout;         /* valid output channels (<= w_rows) */
  int32_t out_nchw_f32; /* 1: `out_nchw` is used instead of `out` */
  int32_t out_reflect;  /* 1: mirror results into out.halo */
  int32_t out_h, out_w; /* interior output extents (validated against the geometry) */
  int32_t fold_w;       /* 1: kw-folded small-channel input; 2: filter column in GEMM-N (see above) */
} dtg_conv_args;
int dtg_conv(const dtg_conv_args* a, const dtg_plane* in, const void* w, int w_rows, int w_cols,
             const float* bias, const dtg_plane* out, float* out_nchw, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Weight gradient (cuDNN wgrad of nn.Conv2d / nn.ConvTranspose2d / nn.Linear backward):
 *   dw[a][b][kh][kw] += sum_{n,oh,ow} p[n,oh,ow,a] * q[n, oh*stride+kh-pad, ow*stride+kw-pad, b]
 * p = low-resolution side (dy of a conv; x of a transposed conv), q = high-resolution side
 * (x of a conv incl. its materialised halo; dy of a transposed conv).  a < pa, b < qb valid channels.
 * tcgen05 GEMM with pixels as the K dimension (both operands MN-major), split-K over pixel tiles into
 * `workspace`, then a deterministic fixed-order reduction that accumulates into dw (fp32, PyTorch
 * layout [a][b][kh][kw]).
 * ------------------------------------------------------------------------------------------- */
typedef struct dtg_wgrad_args {
  int32_t kh, kw, stride, pad;
  int32_t pa, qb; /* valid channel counts */
  /* 0: plain.  1: q is a 16-byte-per-pixel plane (materialised halo >= pad) read kw-folded: GEMM-N =
   * (kw, b), taps = KH.  2: p is a 16-byte-per-pixel plane with a ZERO halo >= KW-1-pad read kw-folded:
   * GEMM-M = (kw', a), taps = KH; q must be zero-padded (no materialised halo).  Stride 1 only.  Used for
   * the weight gradients of the generators' 7x7 head (1) and tail (2). */
  int32_t fold;
} dtg_wgrad_args;
size_t dtg_conv_wgrad_workspace_bytes(const dtg_wgrad_args* a, const dtg_plane* p, const dtg_plane* q);
int dtg_conv_wgrad(const dtg_wgrad_args* a, const dtg_plane* p, const dtg_plane* q, float* dw,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Single-output-channel head convolutions, stride 1 (the last layers of the PatchGAN discriminators,
 * networks.py:337 Conv2d(256,1,4,padding=1) and networks.py:381 Conv2d(128,1,4)) and their gradients: matrix-vector
 * shaped and HBM bound, so they run as coalesced CUDA-core reductions instead of N-padded tensor-core tiles.
 *   w  : the fp32 master weight [1][cin][kh][kw] (PyTorch layout; no packed operand)
 *   fwd: out_nchw [n][1][oh][ow] fp32 = conv(in, w) + bias (zero padding `pad`)
 *   dgrad: dx [n][h][w][>= cin] from channel 0 of the seed-gradient plane dy
 *   wgrad: dw += the weight gradient; per-block partials in `workspace`, summed in a fixed order
 * ------------------------------------------------------------------------------------------- */
int dtg_head1_fwd(const dtg_plane* in, const float* w, const float* bias, int cin, int kh, int kw, int pad,
                  float* out_nchw, int oh, int ow, void* stream);
int dtg_head1_dgrad(const dtg_plane* dy, const float* w, int cin, int kh, int kw, int pad, const dtg_plane* dx, void* stream);
size_t dtg_head1_wgrad_workspace_bytes(int cin, int kh, int kw);
int dtg_head1_wgrad(const dtg_plane* dy, const dtg_plane* in, float* dw, int cin, int kh, int kw, int pad,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused normalisation + affine + activation (+ residual add) forward.  Replaces the ~9-12 ATen
 * kernels per layer of InstanceNorm.forward (modules.py:83-97), CondInstanceNorm.forward
 * (modules.py:120-132), nn.BatchNorm{1,2}d, the following ReLU / LeakyReLU(0.2), the residual
 * `relu(x + out)` (modules.py:186-187,233-234) and nn.ReflectionPad2d of the next conv's input
 * (modules.py:152,172,203,219).
 *   x      : conv output plane (halo 0)
 *   gamma / beta: INSTANCE: [c] params (scale, shift); COND_INSTANCE: [n][c] post-ReLU affine from
 *            dtg_cin_affine_fwd; BATCH: [c] (weight, bias); NONE: ignored
 *   stats  : out, [n][c][2] fp32 (mean, rstd) saved for backward (BATCH: replicated over n)
 *   coef   : out, [n][c][2] fp32 workspace (a, b) with y = act(x*a + b (+ residual))
 *   partial: workspace, dtg_norm_workspace_bytes()
 *   bn_running: BATCH only, [2][c] (running_mean, running_var) updated with `momentum`
 *   phase  : 0 = everything; 1 = statistics partial sums only (BATCH: leaves [c][2] (sum, sumsq)
 *            in `partial` for a cross-GPU all-reduce); 2 = finalise + apply (count_scale = world size)
 * out: plane of the same dtype; if out.halo > 0 the result is mirrored into the halo (reflection).
 * ------------------------------------------------------------------------------------------- */
typedef struct dtg_norm_args {
  int32_t mode; /* DTG_NORM_* */
  int32_t act;  /* DTG_ACT_NONE / RELU / LRELU */
  float eps;
  float momentum;     /* BATCH */
  int32_t phase;      /* 0 / 1 / 2; dtg_norm_bwd also 3 / 4 */
  int32_t world_size; /* BATCH phase 2: statistics were summed over this many ranks */
} dtg_norm_args;
size_t dtg_norm_workspace_bytes(const dtg_plane* x);
int dtg_norm_fwd(const dtg_norm_args* a, const dtg_plane* x, const dtg_plane* residual, const float* gamma,
                 const float* beta, float* bn_running, float* stats, float* coef, float* partial,
                 const dtg_plane* out, void* stream);

/* Backward of dtg_norm_fwd.  g = (fold(dy) + dy2) * act'(y);  dx = rstd*gamma*(g - mean(g) -
 * xhat*sum(g*xhat)/d)  (SURVEY.md 9.1);  d_residual = g.
 *   dy     : gradient w.r.t. the output plane; if dy.halo > 0 the halo holds gradient of the
 *            reflect-padded copy and is folded back (reflection_pad2d_backward)
 *   dy2    : optional second gradient contribution (halo 0), e.g. the residual branch
 *   y      : saved forward output (interior is read for the activation mask)
 *   x      : saved conv output; stats: saved (mean, rstd)
 *   sums   : out [n][c][2] fp32: (sum g, sum g*xhat) per (n,c)  -> COND_INSTANCE d_shift/d_scale;
 *            INSTANCE / BATCH / NONE additionally accumulate d_beta[c] += sum_n, d_gamma[c] += sum_n
 *            (NONE: d_beta only = bias gradient of the preceding conv)
 *   dx     : out plane; halo 0, or (instance / cond-instance / activation-only layers of <= 1024 pixels) a halo whose ring
 *            is left untouched -- dtg_conv's flat-raster DGRAD reads a zero ring of dy as the convolution's padding;
 *            d_res: optional out plane (halo 0)
 *   phase  : as in forward (BATCH: 1 leaves per-channel sums in `partial` for all-reduce);
 *            INSTANCE / NONE only: 4 = everything except the d_gamma / d_beta accumulation, 3 = that accumulation
 *            alone from the sums the phase-4 call left (same arguments; lets the caller issue the small
 *            parameter-gradient reduction on another stream, off the data-gradient chain)
 * ------------------------------------------------------------------------------------------- */
int dtg_norm_bwd(const dtg_norm_args* a, const dtg_plane* dy, const dtg_plane* dy2, const dtg_plane* y,
                 const dtg_plane* x, const float* stats, const float* gamma, float* sums, float* d_gamma,
                 float* d_beta, float* partial, const dtg_plane* dx, const dtg_plane* d_res, void* stream);

/* CondInstanceNorm z-projections (modules.py:111-118,123-124): gamma = relu(Ws z + bs),
 * beta = relu(Wb z + bb), z [n][nz], W [c][nz], out [n][c].  Backward consumes `sums` of
 * dtg_norm_bwd and accumulates dWs, dbs, dWb, dbb and dz (+=). */
int dtg_cin_affine_fwd(const float* z, const float* ws, const float* bs, const float* wb, const float* bb,
                       int n, int c, int nz, float* gamma, float* beta, void* stream);
int dtg_cin_affine_bwd(const float* z, const float* ws, const float* wb, const float* gamma, const float* beta,
                       const float* sums, int n, int c, int nz, float* d_ws, float* d_bs, float* d_wb,
                       float* d_bb, float* d_z, void* stream);

/* ---------------------------------------------------------------------------------------------
 * .npz field preprocessing of the loader (dataloader.py:17-34) for one stack src [b][h][w][cin]
 * (float32, or float64 if is_f64) on the device: channels [0, c), NaN -> 0 / +-inf -> +-max (np.nan_to_num), per
 * (sample, channel) min-max scaling to [-1, 1] in the array's dtype (constant fields -> 0), resize to gh x gw when
 * that differs from h x w -- skimage.transform.resize as its Python-2 releases (<= 0.14) default: bilinear, input
 * coordinate = scale * (o + 0.5) - 0.5, mode 'constant' (samples outside the image are 0), no anti-aliasing, double
 * arithmetic -- and NHWC -> NCHW float32 into dst [b][c][gh][gw].  lohi: scratch of b * c * 2 elements of src's dtype.
 * ------------------------------------------------------------------------------------------- */
int dtg_preprocess_fields(const void* src, int is_f64, int b, int h, int w, int cin, int c, int gh, int gw, float* dst,
                          void* lohi, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Layout conversion at network entry/exit (the reference tensors are fp32 NCHW, dataloader.py:33):
 * dtg_pack_nchw writes channels [c_off, c_off+c) of `dst` from a dense NCHW fp32 tensor (this is
 * also torch.cat on channels, model.py:410,472) and mirrors them into dst.halo (ReflectionPad2d(3),
 * networks.py:159,211).  If tanh_y != NULL (same shape as src) the values are multiplied by
 * 1 - tanh_y^2 first (Tanh backward when src is a gradient w.r.t. a generator output).
 * dtg_unpack_nchw is the inverse for the interior.
 * ------------------------------------------------------------------------------------------- */
int dtg_pack_nchw(const float* src, const float* tanh_y, int n, int c, int h, int w, const dtg_plane* dst, int c_off,
                  int reflect /* 1: mirror into dst.halo; 0: leave the halo untouched (zero padding) */, void* stream);
int dtg_unpack_nchw(const dtg_plane* src, int c_off, int c, float* dst, void* stream);
/* Space-to-depth variant for the 3/6-channel inputs of the stride-2 first layers (networks.py:322,366,446):
 * dst is [n][h/2][w/2][4*cp] (halo 0); pixel (y,x) channel ch is stored in block (y/2, x/2) at channel
 * ((y&1)*2 + (x&1))*cp + c_off + ch.  TMA then reads contiguous rows instead of gathering every other pixel. */
int dtg_pack_nchw_s2d(const float* src, int n, int c, int h, int w, const dtg_plane* dst, int cp, int c_off, void* stream);
/* Weight gradient of such a layer: dtg_conv_wgrad on the space-to-depth plane yields dw2 [cout][4*cp][3][3]; this
 * adds its entries to the PyTorch-layout gradient dw [cout][cin][k][k] (k = 3 or 4) and clears dw2. */
int dtg_s2d_unfold_add(float* dw2, float* dw, int cout, int cin, int k, int cp, void* stream);

/* Sum of up to 3 gradient planes (each optionally with a halo to fold, channel offset c_off[i]),
 * optionally multiplied by tanh'(y) = 1 - y^2 (y dense NCHW fp32, the generator output), written to
 * `out` channels [0,c).  Implements autograd's fan-in accumulation for fake_A / fake_B
 *  fused with Tanh backward (networks.py:188,243).  Also emits the dense NCHW fp32
 * gradient (before the tanh factor) if out_nchw != NULL; add_nchw is an optional dense fp32
 * [n][c][h][w] addend (e.g. dz of the CIN projections added to D_z_B's input gradient). */
int dtg_grad_gather(const dtg_plane* const* srcs, const int* c_off, int nsrc, const float* add_nchw,
                    const float* tanh_y, int c, const dtg_plane* out, float* out_nchw, void* stream);

/* per-channel sum over (n,h,w) of a plane's interior: d_bias[c] += sum, c <= 16 (bias gradients of the network
 * heads).  workspace: >= 8 KB, zero-initialised once by the caller (self-resetting), deterministic two-stage. */
int dtg_channel_sum(const dtg_plane* x, int c, float* d_bias, void* workspace, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused losses (model.py:56-72, 327-334, 432-439, 458-505): each call reduces one term into
 * `scalars` (device fp32 vector) and writes the seed gradient for the backward pass.
 *   dtg_loss_lsgan : scalars[slot_loss] = mean((pred-target)^2), scalars[slot_mean] = mean(pred)
 *                    (slot < 0 skips);  dpred(plane, channel 0) = grad_scale * 2 (pred-target) / count
 *   dtg_loss_l1    : scalars[slot_loss] = mean|a-b|;  da(plane, channels [0,c)) =
 *                    grad_scale * sign(a-b)/count * (tanh_bwd ? 1-a^2 : 1); a, b dense NCHW fp32.
 *                    slot_aux >= 0: scalars[slot_aux] = mean over n of 0.5*sum_c a^2 (kld_std_guss with
 *                    logvar = 0, model.py:45-53,419,490), slots aux+1 / aux+2 = min(a) / max(a).
 * `workspace`: >= 4 KB zero-initialised once by the caller (self-resetting).
 * ------------------------------------------------------------------------------------------- */
int dtg_loss_lsgan(const float* pred, int n, int h, int w, float target, float grad_scale, float* scalars,
                   int slot_loss, int slot_mean, const dtg_plane* dpred, void* workspace, void* stream);
int dtg_loss_l1(const float* a, const float* b, int n, int c, int h, int w, float grad_scale, int tanh_bwd,
                float* scalars, int slot_loss, int slot_aux, const dtg_plane* da, void* workspace, void* stream);

/* Multi-segment form: up to 8 loss terms (each exactly one dtg_loss_lsgan or dtg_loss_l1 call) reduced in ONE launch --
 * e.g. the fake / real pair of a discriminator's D-pass loss (model.py:327-334, 432-434) or the latent discriminator's
 * posterior / prior pair.  kind DTG_LOSS_LSGAN: a = pred [n][1][h][w], slot_aux = the mean(pred) slot; kind DTG_LOSS_L1:
 * a, b [n][c][h][w], slot_aux as in dtg_loss_l1.  `grad`: the segment's seed-gradient plane or NULL.
 * workspace: nseg x 4 KB, zero-initialised once by the caller (self-resetting). */
enum { DTG_LOSS_LSGAN = 0, DTG_LOSS_L1 = 1 };
typedef struct dtg_loss_seg {
  int32_t kind;
  const float* a;
  const float* b;
  int32_t n, c, h, w;
  float target, grad_scale;
  int32_t tanh_bwd, slot_loss, slot_aux;
  const dtg_plane* grad;
} dtg_loss_seg;
int dtg_loss_fused(const dtg_loss_seg* segs, int nseg, float* scalars, void* workspace, void* stream);

/* ---------------------------------------------------------------------------------------------
 * evaluate.py:39-148 variational_ubo: the objective around G_A_B and the update of q(z).
 * dtg_ubo_laplace: scalars[slot_logp] = (1/n) sum over everything of log_prob_laplace(real, fake, logvar_b)
 *   (model.py:24-28; logvar_b is [1][c][h][w], broadcast over n; evaluate.py:93-94) and, if dfake is given, the seed
 *   gradient of mean_n(-log_prob_n) with respect to the generator's PRE-tanh output (fake = tanh(.), networks.py:188).
 * dtg_ubo_latent_step: scalars[slot_kld] = mean_n kld_std_guss(mu, logvar) (model.py:45-53) of the current iterate;
 *   gradients of mean_n(-log_prob_n + kld_n) w.r.t. (mu, logvar) from dz [n][nz] through z = clamp(mu + eps_cur * sd,
 *   -4, 4) (model.py:15-22); one torch.optim.RMSprop step (lr, alpha, eps; square averages sq_mu / sq_logvar,
 *   evaluate.py:65, 119-121) in place; z_out = clamp(mu' + eps_next * sd', -4, 4) (evaluate.py:123).
 * ------------------------------------------------------------------------------------------- */
int dtg_ubo_laplace(const float* fake, const float* real, const float* logvar_b, int n, int c, int h, int w,
                    float* scalars, int slot_logp, const dtg_plane* dfake, void* workspace, void* stream);
int dtg_ubo_latent_step(float* mu, float* logvar, float* sq_mu, float* sq_logvar, const float* eps_cur,
                        const float* eps_next, const float* dz, int n, int nz, float lr, float alpha, float rms_eps,
                        float* z_out, float* scalars, int slot_kld, void* stream);

/* ---------------------------------------------------------------------------------------------
 * clip_grad_norm + Adam over one flat fp32 arena (model.py:447-452,510-515; torch.optim.Adam,
 * betas (beta1, 0.999), eps 1e-8):  dtg_grad_sumsq writes sum(g^2) to *out_sumsq (deterministic
 * two-stage); dtg_adam_clip scales g by min(1, max_norm/(sqrt(sumsq)+1e-6)) IN PLACE (so .grad holds
 * the clipped gradient like the reference) and applies Adam.  hyper (device): [lr, beta1, beta2, eps,
 * max_norm]; step_dev: device int32 step counter (already incremented for this step).
 * grad_scale multiplies g before everything (1/world_size after an all-reduce SUM).
 * ------------------------------------------------------------------------------------------- */
int dtg_grad_sumsq(const float* g, size_t count, float grad_scale, float* out_sumsq, void* workspace, void* stream);
int dtg_adam_clip(float* p, float* g, float* m, float* v, size_t count, const float* hyper, const float* sumsq,
                  const int32_t* step_dev, float grad_scale, void* stream);
int dtg_step_increment(int32_t* step_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DTG_B200_H_ */
