import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dtg  # noqa
from dtg_b200 import engine, networks, ops, _lib as L
engine.set_precision("bf16")
orig = ops.conv
def conv(x, wp, bias, out, **kw):
    try:
        return orig(x, wp, bias, out, **kw)
    except RuntimeError as e:
        print("FAILED conv: x", (x.n, x.h, x.w, x.c, x.halo), "out", None if out is None else (out.n, out.h, out.w, out.c, out.halo), "wp", tuple(wp.shape), kw)
        raise
ops.conv = conv
for n in (2, 80):
    net = networks.ResnetGenerator(3, 3, 32).cuda()
    x = torch.randn(n, 3, 64, 64, device="cuda", requires_grad=True)
    y = net(x)
    y.sum().backward()
    torch.cuda.synchronize()
    print("n", n, "ok", float(x.grad.abs().mean()))
