"""Micro-driver: fused norm forward / backward on one shape, CUDA-graph timed.
usage: python tools/prof_norm.py [c=64] [h=64] [halo=0] [residual=0] [reps=10] [n=80] [in|cin]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dtg  # noqa
from dtg_b200 import _lib as L, ops

c = int(sys.argv[1]) if len(sys.argv) > 1 else 64
h = int(sys.argv[2]) if len(sys.argv) > 2 else 64
halo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
residual = int(sys.argv[4]) if len(sys.argv) > 4 else 0
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 10
N = int(sys.argv[6]) if len(sys.argv) > 6 else 80
mode = {"in": L.NORM_INSTANCE, "cin": L.NORM_COND_INSTANCE}[sys.argv[7] if len(sys.argv) > 7 else "in"]
dt = torch.bfloat16
x = ops.PlaneT(N, h, h, c, 0, dt); x.t.normal_()
out = ops.PlaneT(N, h, h, c, halo, dt)
res = ops.PlaneT(N, h, h, c, halo, dt) if residual else None
if res is not None:
    res.t.normal_()
st = ops.NormState(x)
gshape = (N, c) if mode == L.NORM_COND_INSTANCE else (c,)
gamma, beta = torch.rand(*gshape, device="cuda") + 0.5, torch.randn(*gshape, device="cuda")
dy = ops.PlaneT(N, h, h, c, halo, dt); dy.t.normal_()
dy2 = ops.PlaneT(N, h, h, c, 0, dt) if residual else None
dx = ops.PlaneT(N, h, h, c, 0, dt)
dres = ops.PlaneT(N, h, h, c, 0, dt) if residual else None
dg, db = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
fwd = lambda: ops.norm_fwd(x, out, st, mode=mode, act=L.ACT_RELU, gamma=gamma, beta=beta, residual=res)
bwd = lambda: ops.norm_bwd(dy, dx, st, mode=mode, act=L.ACT_RELU, y=out, x=x, gamma=gamma, dy2=dy2, d_res=dres,
                           d_gamma=dg if mode == L.NORM_INSTANCE else None, d_beta=db if mode == L.NORM_INSTANCE else None)
nbytes = x.t.numel() * 2
for name, f, passes in (("fwd", fwd, 2 + (1 if residual else 0)), ("bwd", bwd, 4 + (2 if residual else 0))):
    f(); f(); torch.cuda.synchronize()
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for _ in range(reps):
            f()
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(); g.replay(); e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print("norm %s [%d,%d,%d,%d] halo %d res %d%s: %.1f us  algorithmic %.0f MB -> %.0f GB/s" %
          (name, N, h, h, c, halo, residual, " (norm_impl %s)" % os.environ.get("DTG_NORM_IMPL", "default"), us, passes * nbytes / 1e6, passes * nbytes / us / 1e3))
if os.environ.get("PROF_KERNELS"):
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            fwd(); bwd()
        torch.cuda.synchronize()
    for e in prof.key_averages():
        if e.device_time_total > 0:
            print("  %-70s n=%d avg %.1f us" % (e.key[:70], e.count, e.device_time_total / e.count))
