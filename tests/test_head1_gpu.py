"""dtg_head1_fwd / dgrad / wgrad (single-output-channel head convolutions, networks.py:337,381) against torch conv2d in
fp64 on the same (bf16- or tf32-rounded) operands.  Tolerance: fp32 accumulation of <= 4096 products, 2e-3 relative."""
import pytest
import torch
import torch.nn.functional as F

import dtg  # noqa: F401
from dtg_b200 import ops

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-12))


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("n,cin,h,k,pad", [(6, 256, 14, 4, 1), (5, 128, 4, 4, 0), (3, 64, 9, 3, 1)])
def test_head1_matches_torch(dt, n, cin, h, k, pad):
    g = torch.Generator(device="cuda").manual_seed(0)
    oh = h + 2 * pad - k + 1
    x = ops.PlaneT(n, h, h, ops.cpad(cin, dt), 0, dt)
    x.t.copy_(torch.randn(x.t.shape, device="cuda", generator=g))
    x.t[..., cin:] = 0
    w = (torch.randn(1, cin, k, k, device="cuda", generator=g) * 0.05).contiguous()
    bias = torch.randn(1, device="cuda", generator=g)
    xr = x.t[..., :cin].permute(0, 3, 1, 2).double()
    # forward
    out = torch.zeros(n, 1, oh, oh, device="cuda")
    ops.head1_fwd(x, w, bias, out, pad)
    ref = F.conv2d(xr, w.double(), bias.double(), padding=pad)
    assert _rel(out, ref) < 2e-3
    # gradients: seed plane holds dy in channel 0
    dy = ops.PlaneT(n, oh, oh, 16, 0, dt)
    dyv = torch.randn(n, oh, oh, device="cuda", generator=g)
    dy.t[..., 0] = dyv
    dyr = dy.t[..., 0].double().unsqueeze(1)
    dx = ops.PlaneT(n, h, h, ops.cpad(cin, dt), 0, dt)
    dx.t.fill_(7.0)
    ops.head1_dgrad(dy, w, dx, pad)
    ref_dx = torch.nn.grad.conv2d_input(xr.shape, w.double(), dyr, padding=pad)
    assert _rel(dx.t[..., :cin].permute(0, 3, 1, 2), ref_dx) < (1e-2 if dt == torch.bfloat16 else 2e-3)   # bf16 store
    dw = torch.full((1, cin, k, k), 0.5, device="cuda")
    ops.head1_wgrad(dy, x, dw, pad)
    ref_dw = torch.nn.grad.conv2d_weight(xr, w.shape, dyr, padding=pad) + 0.5      # accumulates into dw
    assert _rel(dw, ref_dw) < 2e-3
    dw2 = torch.full((1, cin, k, k), 0.5, device="cuda")
    ops.head1_wgrad(dy, x, dw2, pad)
    assert torch.equal(dw, dw2)          # fixed-order reduction: bitwise reproducible
