"""Micro-driver for ncu: one conv configuration through the C ABI, a few launches.
usage: python tools/prof_conv.py <case> [reps]   case in {res, res_dgrad, c3a, c7in, c7out, db3, res_wgrad, c7in_wgrad}"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dtg  # noqa
from dtg_b200 import _lib as L, ops

case = sys.argv[1] if len(sys.argv) > 1 else "res"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
N = 80
dt = torch.bfloat16
g = torch.Generator().manual_seed(0)
CASES = {  # cin, cout, h, k, s, pad, halo, mode
    "res": (128, 128, 32, 3, 1, 1, 1, "fwd"), "res_dgrad": (128, 128, 32, 3, 1, 1, 1, "dgrad"),
    "c3a": (32, 64, 64, 3, 1, 1, 0, "fwd"), "c3b": (64, 32, 64, 3, 1, 1, 0, "fwd"),
    "c7in": (3, 32, 64, 7, 1, 3, 3, "fwd"), "c7out": (32, 3, 64, 7, 1, 3, 0, "fwd"),
    "db3": (256, 256, 15, 4, 1, 1, 0, "fwd"), "down": (64, 128, 64, 3, 2, 1, 0, "fwd"),
    "res_wgrad": (128, 128, 32, 3, 1, 1, 1, "wgrad"), "c7in_wgrad": (3, 32, 64, 7, 1, 3, 3, "wgrad"),
    "c3b_wgrad": (64, 32, 64, 3, 1, 1, 0, "wgrad"), "c7out_dgrad": (32, 3, 64, 7, 1, 3, 0, "dgrad"),
    "db0": (3, 64, 64, 4, 2, 1, 0, "fwd", 160), "db0_wgrad": (3, 64, 64, 4, 2, 1, 0, "wgrad", 160),
    "da0": (3, 32, 64, 3, 2, 1, 0, "fwd", 160), "db4": (256, 1, 14, 4, 1, 1, 0, "fwd", 160),
    "c7in_dgrad": (3, 32, 64, 7, 1, 3, 3, "dgrad"), "c7out_wgrad": (32, 3, 64, 7, 1, 3, 0, "wgrad"), "c3a_wgrad": (32, 64, 64, 3, 1, 1, 0, "wgrad"),
}
cin, cout, h, k, s, pad, halo, mode = CASES[case][:8]
N = CASES[case][8] if len(CASES[case]) > 8 else N
oh = (h + 2 * pad - k) // s + 1
x = ops.PlaneT(N, h, h, ops.cpad(cin, dt), halo, dt); x.t.normal_()
w = (torch.randn(cout, cin, k, k, generator=g) * 0.05).cuda()
if mode == "fwd" and os.environ.get("PROF_HEAD"):     # dense fp32 NCHW head + tanh; PROF_HEAD=2: filter column in GEMM-N
    kwn = os.environ["PROF_HEAD"] == "2"
    wp = ops.pack_conv_weight(w, dt, "fwd_kwn" if kwn else "fwd")
    y = torch.empty(N, cout, oh, oh, device="cuda")
    bias = torch.zeros(cout, device="cuda")
    f = lambda: ops.conv(x, wp, bias, None, kh=k, kw=k, stride=s, pad=pad, cout=cout, out_h=oh, out_w=oh, act=L.ACT_TANH,
                         out_nchw=y, fold_w=2 if kwn else 0)
elif mode == "fwd":
    wp = ops.pack_conv_weight(w, dt, "fwd")
    out = ops.PlaneT(N, oh, oh, ops.cpad(cout, dt), 0, dt)
    bias = torch.zeros(cout, device="cuda") if os.environ.get("PROF_BIAS") else None
    f = lambda: ops.conv(x, wp, bias, out, kh=k, kw=k, stride=s, pad=pad, cout=cout, out_h=oh, out_w=oh,
                         act=int(os.environ.get("PROF_ACT", "0")))
elif mode == "dgrad":
    wp = ops.pack_conv_weight(w, dt, "dgrad")
    flat = bool(os.environ.get("PROF_FLAT")) and halo > 0      # dy with a zero halo ring == ring: conv_patch2.cu flat-raster mode
    dy = ops.PlaneT(N, oh, oh, ops.cpad(cout, dt), halo if flat else 0, dt)
    dy.interior().normal_() if flat else dy.t.normal_()
    dx = ops.PlaneT(N, h, h, ops.cpad(cin, dt), halo, dt)
    f = lambda: ops.conv(dy, wp, None, dx, mode=L.CONV_DGRAD, kh=k, kw=k, stride=s, pad=pad, ring=halo, cout=cin, out_h=h, out_w=h)
else:
    dy = ops.PlaneT(N, oh, oh, ops.cpad(cout, dt), 0, dt); dy.t.normal_()
    dw = torch.zeros(cout, cin, k, k, device="cuda")
    f = lambda: ops.conv_wgrad(dy, x, dw, kh=k, kw=k, stride=s, pad=pad, pa=cout, qb=cin)
f(); f(); torch.cuda.synchronize()
# `reps` back-to-back launches captured in one CUDA graph: pure GPU time (no host launch gaps)
side = torch.cuda.Stream()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=side):
    for i in range(reps):
        f()
g.replay()
ts = []
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(); g.replay(); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3 / reps)
fl = 2.0 * N * oh * oh * cin * cout * k * k
print(case, "us per call (graph of %d):" % reps, ["%.1f" % t for t in ts], "TFLOP/s best %.1f" % (fl / min(ts) / 1e6))
if os.environ.get("PROF_KERNELS"):
    # per-kernel device durations (CUPTI activity records through torch.profiler; no replay, warm caches)
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            f()
        torch.cuda.synchronize()
    for e in prof.key_averages():
        if e.device_time_total > 0:
            print("  %-70s n=%d avg %.1f us" % (e.key[:70], e.count, e.device_time_total / e.count))
