"""GPU parity of the tcgen05 split-K weight-gradient GEMM against torch autograd (fp64 reference on
bf16-/tf32-representable operands).  The accumulator is fp32 and the output fp32, so the only error
is summation order: tolerance 2e-3 of the gradient's max magnitude (tf32 operands are pre-truncated)."""
import pytest
import torch
import torch.nn.functional as F

import dtg  # noqa: F401
from dtg_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _q(t, dtype):
    if dtype == torch.bfloat16:
        return t.to(torch.bfloat16).float()
    return ((t.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)   # round-to-nearest tf32


CASES = [
    # n, cin, cout, h, k, s, pad, halo
    (3, 128, 128, 32, 3, 1, 1, 1),     # res-block conv on a reflect-padded plane
    (2, 64, 128, 64, 3, 2, 1, 0),      # generator downsample (stride 2, parity maps)
    (2, 32, 64, 64, 3, 1, 1, 0),
    (2, 64, 32, 64, 3, 1, 1, 0),
    (2, 3, 32, 64, 7, 1, 3, 3),        # 7x7 head, 49 taps
    (2, 32, 3, 64, 7, 1, 3, 0),        # 7x7 tail, 3 output channels
    (2, 3, 64, 64, 4, 2, 1, 0),
    (2, 128, 256, 16, 4, 1, 1, 0),     # 256 output channels: two M blocks
    (2, 256, 256, 15, 4, 1, 1, 0),
    (2, 256, 1, 14, 4, 1, 1, 0),
    (5, 128, 128, 8, 3, 2, 1, 0),
    (70, 16, 64, 1, 1, 1, 0, 0),       # Linear
    (6, 256, 256, 4, 4, 1, 0, 0),
]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("case", CASES)
def test_conv_wgrad(case, dtype):
    n, cin, cout, h, k, s, pad, halo = case
    g = torch.Generator().manual_seed(hash(case) % 1000)
    oh = (h + 2 * pad - k) // s + 1
    x = _q(torch.randn(n, cin, h, h, generator=g), dtype).to(DEV)
    dy = _q(torch.randn(n, cout, oh, oh, generator=g), dtype).to(DEV)
    xq = ops.PlaneT.from_nchw(x, halo=halo, dtype=dtype)
    dyp = ops.PlaneT.from_nchw(dy, dtype=dtype)
    dw = torch.full((cout, cin, k, k), 0.5, device=DEV)       # accumulate semantics (+=)
    ops.conv_wgrad(dyp, xq, dw, kh=k, kw=k, stride=s, pad=pad, pa=cout, qb=cin)
    xin = F.pad(x, (pad,) * 4, mode="reflect") if halo else F.pad(x, (pad,) * 4)
    wt = torch.zeros(cout, cin, k, k, device=DEV, dtype=torch.float64, requires_grad=True)
    F.conv2d(xin.double(), wt, stride=s).backward(dy.double())
    ref = wt.grad.float() + 0.5
    err = float((dw - ref).abs().max() / ref.abs().max())
    assert err < 2e-3, (case, err)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_conv_wgrad_reads_the_interior_of_a_haloed_dy(dtype):
    """the output gradient may live in a plane with a (zero) halo ring (the flat-raster dgrad of conv_patch2.cu wants one):
    the weight gradient views its interior and is bit-identical to the halo-free call"""
    g = torch.Generator().manual_seed(4)
    n, c, h = 3, 128, 32
    x = _q(torch.randn(n, c, h, h, generator=g), dtype).to(DEV)
    dy = _q(torch.randn(n, c, h, h, generator=g), dtype).to(DEV)
    xq = ops.PlaneT.from_nchw(x, halo=1, dtype=dtype)
    dw0, dw1 = torch.zeros(c, c, 3, 3, device=DEV), torch.zeros(c, c, 3, 3, device=DEV)
    ops.conv_wgrad(ops.PlaneT.from_nchw(dy, dtype=dtype), xq, dw0, kh=3, kw=3, stride=1, pad=1, pa=c, qb=c)
    dyh = ops.PlaneT.from_nchw(dy, halo=1, dtype=dtype, reflect=False)
    dyh.t[:, 0].fill_(7.0)          # the ring must be ignored, whatever it holds
    dyh.t[:, :, 0].fill_(-3.0)
    ops.conv_wgrad(dyh, xq, dw1, kh=3, kw=3, stride=1, pad=1, pa=c, qb=c)
    assert torch.equal(dw0, dw1)


def test_conv_transpose_wgrad():
    """ConvTranspose2d weight [ci, co, kh, kw]: p = x (low-res, ci), q = dy (high-res, co)."""
    dtype = torch.bfloat16
    g = torch.Generator().manual_seed(3)
    x = _q(torch.randn(2, 128, 32, 32, generator=g), dtype).to(DEV)
    dy = _q(torch.randn(2, 64, 64, 64, generator=g), dtype).to(DEV)
    dw = torch.zeros(128, 64, 3, 3, device=DEV)
    ops.conv_wgrad(ops.PlaneT.from_nchw(x, dtype=dtype), ops.PlaneT.from_nchw(dy, dtype=dtype), dw,
                   kh=3, kw=3, stride=2, pad=1, pa=128, qb=64)
    wt = torch.zeros(128, 64, 3, 3, device=DEV, dtype=torch.float64, requires_grad=True)
    F.conv_transpose2d(x.double(), wt, stride=2, padding=1, output_padding=1).backward(dy.double())
    ref = wt.grad.float()
    assert float((dw - ref).abs().max() / ref.abs().max()) < 2e-3


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("k,h,w", [(7, 64, 64), (5, 24, 40)])
def test_conv_wgrad_fold_q(k, h, w, dtype):
    """fold=1: the small-channel INPUT (16-byte pixels, materialised reflect halo) is read kw-folded (7x7 head)."""
    n, cin, cout, pad = 2, 3, 32, (k - 1) // 2
    g = torch.Generator().manual_seed(7 + k)
    x = _q(torch.randn(n, cin, h, w, generator=g), dtype).to(DEV)
    dy = _q(torch.randn(n, cout, h, w, generator=g), dtype).to(DEV)
    xq = ops.PlaneT.from_nchw(x, halo=pad, dtype=dtype, c_store=ops.fold_channels(dtype))
    dyp = ops.PlaneT.from_nchw(dy, dtype=dtype)
    dw = torch.full((cout, cin, k, k), 0.25, device=DEV)
    ops.conv_wgrad(dyp, xq, dw, kh=k, kw=k, stride=1, pad=pad, pa=cout, qb=cin, fold=1)
    wt = torch.zeros(cout, cin, k, k, device=DEV, dtype=torch.float64, requires_grad=True)
    F.conv2d(F.pad(x, (pad,) * 4, mode="reflect").double(), wt).backward(dy.double())
    ref = wt.grad.float() + 0.25
    assert float((dw - ref).abs().max() / ref.abs().max()) < 2e-3


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("k,h,w", [(7, 64, 64), (5, 24, 40)])
def test_conv_wgrad_fold_p(k, h, w, dtype):
    """fold=2: the small-channel OUTPUT GRADIENT (16-byte pixels, zero halo) is read kw-folded (7x7 tail)."""
    n, cin, cout, pad = 2, 32, 3, (k - 1) // 2
    g = torch.Generator().manual_seed(11 + k)
    x = _q(torch.randn(n, cin, h, w, generator=g), dtype).to(DEV)
    dy = _q(torch.randn(n, cout, h, w, generator=g), dtype).to(DEV)
    xq = ops.PlaneT.from_nchw(x, dtype=dtype)
    dyp = ops.PlaneT.from_nchw(dy, halo=pad, dtype=dtype, c_store=ops.fold_channels(dtype), reflect=False)
    dw = torch.zeros(cout, cin, k, k, device=DEV)
    ops.conv_wgrad(dyp, xq, dw, kh=k, kw=k, stride=1, pad=pad, pa=cout, qb=cin, fold=2)
    wt = torch.zeros(cout, cin, k, k, device=DEV, dtype=torch.float64, requires_grad=True)
    F.conv2d(x.double(), wt, padding=pad).backward(dy.double())
    ref = wt.grad.float()
    assert float((dw - ref).abs().max() / ref.abs().max()) < 2e-3
