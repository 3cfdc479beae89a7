"""GPU parity of the tcgen05 implicit-GEMM convolution (through the C ABI) against torch's fp32
convolution on the same bf16-/tf32-representable operands.  Tolerances: bf16 planes store 8
significant bits -> |err| <= 2^-8 * |y| + accumulation noise; TF32 mode compares at 2e-3 relative
(10-bit mantissa operands, fp32 accumulate)."""
import pytest
import torch
import torch.nn.functional as F

import dtg  # noqa: F401
from dtg_b200 import _lib as L, ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _q(t, dtype):
    """round to what the plane stores (bf16) / what tcgen05 kind::tf32 consumes (10-bit mantissa, truncation)."""
    if dtype == torch.bfloat16:
        return t.to(torch.bfloat16).float()
    return ((t.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)   # round-to-nearest tf32


def _tol(dtype):
    return 1.2e-2 if dtype == torch.bfloat16 else 2e-3


def _relerr(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-6))


CASES = [
    # n, cin, cout, h, w, k, stride, pad, halo(reflect)
    (2, 64, 128, 32, 32, 3, 1, 1, 0),
    (2, 128, 128, 32, 32, 3, 1, 1, 1),     # res-block conv, reflect halo
    (3, 64, 128, 64, 64, 3, 2, 1, 0),      # generator downsample
    (2, 32, 64, 64, 64, 3, 1, 1, 0),       # cin < one 128-byte chunk
    (2, 3, 32, 64, 64, 7, 1, 3, 3),        # 7x7 head, reflect halo 3
    (2, 3, 64, 64, 64, 4, 2, 1, 0),        # PatchGAN first layer
    (2, 128, 256, 16, 16, 4, 1, 1, 0),     # PatchGAN 16 -> 15
    (2, 256, 256, 15, 15, 4, 1, 1, 0),     # 15 -> 14
    (5, 128, 128, 8, 8, 3, 2, 1, 0),       # D_A / E tail (several images per tile)
    (130, 16, 64, 1, 1, 1, 1, 0, 0),       # Linear as 1x1 conv on [N,1,1,C]
    (4, 256, 256, 4, 4, 4, 1, 0, 0),       # encoder 4x4 valid conv -> 1x1
    (2, 64, 32, 64, 64, 3, 1, 1, 0),       # patch-resident kernel, one full 128-byte chunk
    (3, 16, 32, 13, 21, 3, 1, 1, 0),       # patch-resident kernel, ragged tile edges
    (2, 32, 16, 40, 24, 5, 1, 2, 2),       # 5x5 reflect halo 2
]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("case", CASES)
def test_conv_fwd(case, dtype):
    n, cin, cout, h, w, k, s, pad, halo = case
    g = torch.Generator(device="cpu").manual_seed(hash(case) % 1000)
    x = _q(torch.randn(n, cin, h, w, generator=g), dtype).to(DEV)
    wt = _q(torch.randn(cout, cin, k, k, generator=g) * 0.1, dtype).to(DEV)
    b = torch.randn(cout, generator=g).to(DEV)
    xp = ops.PlaneT.from_nchw(x, halo=halo, dtype=dtype)
    wp = ops.pack_conv_weight(wt, dtype, "fwd")
    oh = (h + 2 * pad - k) // s + 1
    ow = (w + 2 * pad - k) // s + 1
    out = ops.PlaneT(n, oh, ow, ops.cpad(cout, dtype), 0, dtype)
    ops.conv(xp, wp, b, out, kh=k, kw=k, stride=s, pad=pad, act=L.ACT_LRELU, cout=cout, out_h=oh, out_w=ow)
    xin = F.pad(x, (pad,) * 4, mode="reflect") if halo else F.pad(x, (pad,) * 4)
    with torch.backends.cudnn.flags(allow_tf32=False):
        torch.backends.cuda.matmul.allow_tf32 = False
        ref = F.leaky_relu(F.conv2d(xin.double(), wt.double(), b.double(), stride=s), 0.2).float()
    got = out.to_nchw(cout)
    assert _relerr(got, ref) < _tol(dtype), (case, _relerr(got, ref))
    if ops.cpad(cout, dtype) > cout:   # padded channels must be exact zeros
        assert float(out.interior()[..., cout:].abs().max()) == 0.0


def test_conv_fwd_nchw_tanh_and_reflect_out():
    dtype = torch.bfloat16
    g = torch.Generator().manual_seed(5)
    x = _q(torch.randn(2, 32, 64, 64, generator=g), dtype).to(DEV)
    wt = _q(torch.randn(3, 32, 7, 7, generator=g) * 0.05, dtype).to(DEV)
    b = torch.randn(3, generator=g).to(DEV) * 0.1
    xp = ops.PlaneT.from_nchw(x, dtype=dtype)
    wp = ops.pack_conv_weight(wt, dtype, "fwd")
    y = torch.empty(2, 3, 64, 64, device=DEV)
    ops.conv(xp, wp, b, None, kh=7, kw=7, pad=3, act=L.ACT_TANH, cout=3, out_h=64, out_w=64, out_nchw=y)
    ref = torch.tanh(F.conv2d(x.double(), wt.double(), b.double(), padding=3)).float()
    assert _relerr(y, ref) < 2e-3
    # reflect-mirrored output halo
    wt2 = _q(torch.randn(64, 32, 3, 3, generator=g) * 0.1, dtype).to(DEV)
    wp2 = ops.pack_conv_weight(wt2, dtype, "fwd")
    out = ops.PlaneT(2, 64, 64, 64, 1, dtype)
    ops.conv(xp, wp2, None, out, kh=3, kw=3, pad=1, act=L.ACT_RELU, cout=64, out_h=64, out_w=64, out_reflect=True)
    ref2 = F.pad(F.relu(F.conv2d(x.double(), wt2.double(), padding=1)), (1,) * 4, mode="reflect").float()
    got2 = out.t.permute(0, 3, 1, 2).float()
    assert _relerr(got2, ref2) < 1.2e-2


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("n,cin,cout,h,w,k", [(2, 32, 3, 64, 64, 7), (3, 32, 1, 128, 128, 7), (2, 32, 3, 32, 32, 7),
                                              (160, 32, 3, 64, 64, 7), (2, 16, 4, 24, 16, 5), (1, 32, 3, 33, 64, 7)])
def test_conv_tail_filter_column_in_gemm_n(n, cin, cout, h, w, k, dtype):
    """dtg_conv fold_w = 2 (conv_tail7.cu): the generators' 7x7 tail + tanh (networks.py:187-188, 242-243) with (kw, cout)
    in GEMM-N and a shift-add epilogue, against torch's fp64 convolution on the same representable operands; 160 images =
    more tiles than SMs (persistent loop, 4 TMEM buffers wrap), 33 rows = a ragged last tile, 24x16 = 8 rows per tile"""
    g = torch.Generator().manual_seed(11)
    x = _q(torch.randn(n, cin, h, w, generator=g), dtype).to(DEV)
    wt = _q(torch.randn(cout, cin, k, k, generator=g) * 0.05, dtype).to(DEV)
    b = torch.randn(cout, generator=g).to(DEV) * 0.1
    xp = ops.PlaneT.from_nchw(x, dtype=dtype)
    wp = ops.pack_conv_weight(wt, dtype, "fwd_kwn")
    assert tuple(wp.shape)[:2] == (k, 32)
    y = torch.full((n, cout, h, w), float("nan"), device=DEV)
    run = lambda: ops.conv(xp, wp, b, None, kh=k, kw=k, pad=k // 2, act=L.ACT_TANH, cout=cout, out_h=h, out_w=w, out_nchw=y,
                           fold_w=2)
    if not ops.tail_kwn_eligible(xp.c, k, cout, w, dtype):       # fp32 at width 128: two 114 KB patch stages do not fit
        assert dtype == torch.float32 and w == 128
        with pytest.raises(RuntimeError, match="not eligible"):
            run()
        return
    run()
    ref = torch.tanh(F.conv2d(x.double(), wt.double(), b.double(), padding=k // 2)).float()
    assert torch.isfinite(y).all()
    assert _relerr(y, ref) < 2e-3
    # identical (up to summation order) to the ordinary tap-per-MMA mapping
    y2 = torch.empty_like(y)
    ops.conv(xp, ops.pack_conv_weight(wt, dtype, "fwd"), b, None, kh=k, kw=k, pad=k // 2, act=L.ACT_TANH, cout=cout, out_h=h,
             out_w=w, out_nchw=y2)
    assert _relerr(y, y2) < 1e-4


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("n,cin,cout,h,w,k", [(2, 3, 32, 64, 64, 7), (160, 3, 32, 64, 64, 7), (3, 1, 32, 33, 32, 7), (2, 4, 16, 24, 16, 5)])
def test_conv_head_dgrad_filter_column_in_gemm_n(n, cin, cout, h, w, k, dtype):
    """dtg_conv fold_w = 2, DGRAD (conv_tail7.cu, full mode): gradient of the generators' 7x7 head (networks.py:159-160,
    211-212) w.r.t. its reflect-PADDED input -- every pixel of the haloed 16-byte-pixel plane -- against the transposed
    convolution in fp64, and against the tap-per-MMA mapping"""
    g = torch.Generator().manual_seed(13)
    pad = k // 2
    dy = _q(torch.randn(n, cout, h, w, generator=g), dtype).to(DEV)
    wt = _q(torch.randn(cout, cin, k, k, generator=g) * 0.05, dtype).to(DEV)
    dyp = ops.PlaneT.from_nchw(dy, dtype=dtype)
    cs = ops.cpad_small(cin, dtype)
    if not ops.tail_kwn_eligible(dyp.c, k, cin, w, dtype):
        pytest.skip("patch stages do not fit")
    wp = ops.pack_conv_weight(wt, dtype, "dgrad_kwn")
    dx = ops.PlaneT(n, h, w, cs, pad, dtype)
    dx.t.fill_(float("nan"))
    ops.conv(dyp, wp, None, dx, mode=L.CONV_DGRAD, kh=k, kw=k, pad=pad, ring=pad, cout=cin, out_h=h, out_w=w, fold_w=2)
    ref = F.conv_transpose2d(dy.double(), wt.double()).float()                     # [n, cin, h + k - 1, w + k - 1]
    got = dx.t.permute(0, 3, 1, 2).float()
    assert torch.isfinite(got).all()
    assert _relerr(got[:, :cin], ref) < _tol(dtype)
    assert float(got[:, cin:].abs().max()) == 0.0 if cs > cin else True
    dx2 = ops.PlaneT(n, h, w, cs, pad, dtype)
    ops.conv(dyp, ops.pack_conv_weight(wt, dtype, "dgrad"), None, dx2, mode=L.CONV_DGRAD, kh=k, kw=k, pad=pad, ring=pad, cout=cin,
             out_h=h, out_w=w)
    assert _relerr(got[:, :cin], dx2.t.permute(0, 3, 1, 2).float()[:, :cin]) < (1e-2 if dtype == torch.bfloat16 else 1e-3)   # both outputs are rounded to the plane's precision


DGRAD_CASES = [
    # n, cin, cout, h(in), k, s, pad, ring
    (2, 128, 128, 32, 3, 1, 1, 0),
    (2, 128, 128, 32, 3, 1, 1, 1),     # gradient w.r.t. the reflect-padded input (34x34)
    (2, 64, 128, 64, 3, 2, 1, 0),      # stride-2 dgrad: 4 parity phases
    (2, 64, 128, 32, 4, 2, 1, 0),
    (2, 3, 32, 64, 7, 1, 3, 3),
    (2, 32, 3, 64, 7, 1, 3, 0),        # dgrad of the generator's 7x7 tail (3-channel dy)
    (2, 64, 32, 64, 3, 1, 1, 0),
    (2, 128, 256, 16, 4, 1, 1, 0),
    (3, 256, 256, 4, 4, 1, 0, 0),
]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("case", DGRAD_CASES)
def test_conv_dgrad(case, dtype):
    n, cin, cout, h, k, s, pad, ring = case
    g = torch.Generator().manual_seed(hash(case) % 1000)
    oh = (h + 2 * pad - k) // s + 1
    dy = _q(torch.randn(n, cout, oh, oh, generator=g), dtype).to(DEV)
    wt = _q(torch.randn(cout, cin, k, k, generator=g) * 0.1, dtype).to(DEV)
    dyp = ops.PlaneT.from_nchw(dy, dtype=dtype)
    wp = ops.pack_conv_weight(wt, dtype, "dgrad")
    dx = ops.PlaneT(n, h, h, ops.cpad(cin, dtype), ring, dtype)
    ops.conv(dyp, wp, None, dx, mode=L.CONV_DGRAD, kh=k, kw=k, stride=s, pad=pad, ring=ring, cout=cin, out_h=h, out_w=h)
    # reference: gradient of conv w.r.t. an input padded by `ring` (conv pad reduced accordingly)
    xin = torch.zeros(n, cin, h + 2 * ring, h + 2 * ring, device=DEV, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(xin, wt.double(), stride=s, padding=pad - ring)
    y.backward(dy.double())
    ref = xin.grad.float()
    got = dx.t.permute(0, 3, 1, 2).float()[:, :cin]
    assert _relerr(got, ref) < _tol(dtype), (case, _relerr(got, ref))


@pytest.mark.parametrize("n,c,h", [(2, 128, 32), (80, 128, 32), (4, 128, 16), (4, 64, 20)])
def test_conv_dgrad_flat_raster(n, c, h):
    """dtg_conv DGRAD with a dy plane whose zero halo equals the ring (conv_patch2.cu flat mode): the gradient w.r.t. the
    reflect-padded input of the residual-stack convs (modules.py:162,180,211,227), all images tiled as one tall image;
    against autograd in fp64 and against the per-tap kernel on a halo-free dy"""
    dtype = torch.bfloat16
    g = torch.Generator().manual_seed(17)
    dy = _q(torch.randn(n, c, h, h, generator=g), dtype).to(DEV)
    wt = _q(torch.randn(c, c, 3, 3, generator=g) * 0.05, dtype).to(DEV)
    wp = ops.pack_conv_weight(wt, dtype, "dgrad")
    assert ops.flat_dgrad_eligible(n, h, h, c, c, 1, dtype)
    dyh = ops.PlaneT.from_nchw(dy, halo=1, dtype=dtype, reflect=False)          # zero ring
    assert float(dyh.t[:, 0].abs().max()) == 0.0 and float(dyh.t[:, :, -1].abs().max()) == 0.0
    dx = ops.PlaneT(n, h, h, c, 1, dtype)
    dx.t.fill_(float("nan"))
    ops.conv(dyh, wp, None, dx, mode=L.CONV_DGRAD, kh=3, kw=3, pad=1, ring=1, cout=c, out_h=h, out_w=h)
    xin = torch.zeros(n, c, h + 2, h + 2, device=DEV, dtype=torch.float64, requires_grad=True)
    F.conv2d(xin, wt.double()).backward(dy.double())
    got = dx.t.permute(0, 3, 1, 2).float()
    assert torch.isfinite(got).all()
    assert _relerr(got, xin.grad.float()) < _tol(dtype)
    dx0 = ops.PlaneT(n, h, h, c, 1, dtype)
    ops.conv(ops.PlaneT.from_nchw(dy, dtype=dtype), wp, None, dx0, mode=L.CONV_DGRAD, kh=3, kw=3, pad=1, ring=1, cout=c, out_h=h,
             out_w=h)
    assert _relerr(dx.t.float(), dx0.t.float()) < 1e-2         # same products, different fp32 summation order, bf16 outputs


def test_conv_transpose_forward():
    """ConvTranspose2d(128->64, k3, s2, p1, op1) forward = DGRAD mode with the 'tfwd' packing."""
    dtype = torch.bfloat16
    g = torch.Generator().manual_seed(11)
    x = _q(torch.randn(2, 128, 32, 32, generator=g), dtype).to(DEV)
    wt = _q(torch.randn(128, 64, 3, 3, generator=g) * 0.1, dtype).to(DEV)
    b = torch.randn(64, generator=g).to(DEV)
    xp = ops.PlaneT.from_nchw(x, dtype=dtype)
    wp = ops.pack_conv_weight(wt, dtype, "tfwd")
    out = ops.PlaneT(2, 64, 64, 64, 0, dtype)
    ops.conv(xp, wp, b, out, mode=L.CONV_DGRAD, kh=3, kw=3, stride=2, pad=1, cout=64, out_h=64, out_w=64)
    ref = F.conv_transpose2d(x.double(), wt.double(), b.double(), stride=2, padding=1, output_padding=1).float()
    assert _relerr(out.to_nchw(64), ref) < 1.2e-2


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("k,h,w", [(7, 64, 64), (5, 24, 40)])
def test_conv_fold_fwd_and_dgrad(k, h, w, dtype):
    """kw-folded small-channel 7x7 (5x5) layers: forward on a 16-byte-per-pixel reflect-padded input, and the data
    gradient of a 3-channel tail from a 16-byte-per-pixel dy with a zero halo (overlapping-row TMA views)."""
    n, pad = 2, (k - 1) // 2
    g = torch.Generator().manual_seed(21 + k)
    fc = ops.fold_channels(dtype)
    # forward 3 -> 32
    x = _q(torch.randn(n, 3, h, w, generator=g), dtype).to(DEV)
    wt = _q(torch.randn(32, 3, k, k, generator=g) * 0.1, dtype).to(DEV)
    xp = ops.PlaneT.from_nchw(x, halo=pad, dtype=dtype, c_store=fc)
    out = ops.PlaneT(n, h, w, 32, 0, dtype)
    ops.conv(xp, ops.pack_conv_weight(wt, dtype, "fwd_fold"), None, out, kh=k, kw=k, pad=pad, cout=32, out_h=h, out_w=w,
             fold_w=True)
    ref = F.conv2d(F.pad(x, (pad,) * 4, mode="reflect").double(), wt.double()).float()
    assert _relerr(out.to_nchw(32), ref) < _tol(dtype)
    # dgrad of a 32 -> 3 conv: dy has 3 channels
    dy = _q(torch.randn(n, 3, h, w, generator=g), dtype).to(DEV)
    wt2 = _q(torch.randn(3, 32, k, k, generator=g) * 0.1, dtype).to(DEV)
    dyp = ops.PlaneT.from_nchw(dy, halo=pad, dtype=dtype, c_store=fc, reflect=False)
    dx = ops.PlaneT(n, h, w, 32, 0, dtype)
    ops.conv(dyp, ops.pack_conv_weight(wt2, dtype, "dgrad_fold"), None, dx, mode=L.CONV_DGRAD, kh=k, kw=k, pad=pad, cout=32,
             out_h=h, out_w=w, fold_w=True)
    xin = torch.zeros(n, 32, h, w, device=DEV, dtype=torch.float64, requires_grad=True)
    F.conv2d(xin, wt2.double(), padding=pad).backward(dy.double())
    assert _relerr(dx.to_nchw(32), xin.grad.float()) < _tol(dtype)
