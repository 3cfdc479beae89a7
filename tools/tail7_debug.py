import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import dtg  # noqa
from dtg_b200 import _lib as L, ops
g = torch.Generator().manual_seed(1)
for dtype in (torch.float32, torch.bfloat16):
    for n in (8, 5, 16, 80):
        x = torch.randn(n, 32, 64, 64, generator=g).cuda()
        if dtype == torch.float32:
            x = ((x.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
        else:
            x = x.bfloat16().float()
        wt = (torch.randn(3, 32, 7, 7, generator=g) * 0.05).cuda()
        b = torch.randn(3, generator=g).cuda() * 0.1
        wp = ops.pack_conv_weight(wt, dtype, "fwd_kwn")
        def run(xx):
            xp = ops.PlaneT.from_nchw(xx, dtype=dtype)
            y = torch.full((n, 3, 64, 64), float("nan"), device="cuda")
            ops.conv(xp, wp, b, None, kh=7, kw=7, pad=3, act=L.ACT_TANH, cout=3, out_h=64, out_w=64, out_nchw=y, fold_w=2)
            torch.cuda.synchronize()
            return y
        y = run(x)
        perm = torch.arange(n - 1, -1, -1, device="cuda")
        yp = run(x[perm].contiguous())
        same = torch.equal(yp, y[perm])
        wq = wt if dtype == torch.bfloat16 else ((wt.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
        if dtype == torch.bfloat16: wq = wt.bfloat16().float()
        ref = torch.tanh(F.conv2d(x.double(), wq.double(), b.double(), padding=3)).float()
        err = (y - ref).abs().amax(dim=(1, 2, 3))
        d = (yp - y[perm]).abs().amax(dim=(1, 2, 3))
        print(str(dtype), "n", n, "perm-invariant", same, "max err per sample", [round(float(e), 5) for e in err][:16], "perm diff", [round(float(e), 5) for e in d][:16])
        if not same or float(err.max()) > 5e-3:
            bad = (y - ref).abs()
            i = int(err.argmax())
            rows = bad[i].amax(dim=(0, 2))
            print("   worst sample", i, "bad rows:", [int(r) for r in torch.nonzero(rows > 5e-3).flatten()[:40]])
