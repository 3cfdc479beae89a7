"""Kernel timeline of graph-replayed train_instance steps through torch.profiler (CUPTI activity records): which
kernels overlap, how busy each resource class is, where the step waits.  Analysis only -- numbers taken under a
profiler are never bench values.

usage: python tools/timeline.py [--batch 80] [--out gpurun_out/timeline.json]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import dtg  # noqa: E402,F401
from dtg_b200 import engine, model as dmodel  # noqa: E402
from oracle import step as ostep  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=80)
ap.add_argument("--out", default="gpurun_out/timeline.json")
ap.add_argument("--serial", action="store_true")
args = ap.parse_args()

engine.set_precision("bf16")
opt = argparse.Namespace(**vars(ostep.default_opt()), expr_dir="/tmp", niter_decay=25)
torch.manual_seed(1234)
m = dmodel.AugmentedCycleGAN(opt, testing=True)
m.prepare()
if args.serial:
    m.lanes.enabled = False
dev = [t.cuda() for t in ostep.synthetic_batch(args.batch, seed=4321)]
for _ in range(5):
    m.train_instance(*dev, use_graph=True, report=False)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        m.train_instance(*dev, use_graph=True, report=False)
    torch.cuda.synchronize()
ev = []
streams = {}
try:
    for k in prof.profiler.kineto_results.events():
        if "CUDA" in str(k.device_type()) and k.duration_ns() > 0:
            a = k.start_ns() / 1e3
            ev.append((a, a + k.duration_ns() / 1e3, k.name()))
            streams[(a, k.name())] = k.device_resource_id()
except Exception as exc:       # older / newer kineto bindings: fall back to the python events (no stream ids)
    sys.stderr.write("kineto events unavailable (%s)\n" % exc)
    ev = []
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None:
            ev.append((e.time_range.start, e.time_range.end, e.name))
ev.sort()
if not ev:
    raise SystemExit("no CUDA activity records (CUPTI unavailable?)")
# split into the three replays by the largest gaps
t0, t1 = ev[0][0], max(e[1] for e in ev)
print("records %d, span %.2f ms (3 steps)" % (len(ev), (t1 - t0) / 1e3))


def cls(name):
    n = name
    if "wgrad_reduce" in n:
        return "wgrad_reduce"
    if "wgrad" in n:
        return "wgrad"
    if "igemm" in n or "pconv" in n:
        return "conv"
    if "norm_bwd" in n or "bn_" in n:
        return "norm_bwd"
    if "norm" in n:
        return "norm_fwd"
    return "other"


# take the middle third of the records as "one step"
third = len(ev) // 3
step = ev[third:2 * third]
s0, s1 = step[0][0], max(e[1] for e in step)
print("one step: %.2f ms, %d kernels" % ((s1 - s0) / 1e3, len(step)))
# sweep: time with k kernels active; busy time per class; time with any tensor kernel active
pts = []
for a, b, n in step:
    pts.append((a, 1, cls(n)))
    pts.append((b, -1, cls(n)))
pts.sort(key=lambda p: (p[0], p[1]))
active = {}
hist = {}
cls_busy = {}
tensor_busy = 0.0
last = s0
for t, d, c in pts:
    dt = t - last
    if dt > 0:
        k = sum(active.values())
        hist[k] = hist.get(k, 0.0) + dt
        for cc, v in active.items():
            if v > 0:
                cls_busy[cc] = cls_busy.get(cc, 0.0) + dt
        if active.get("conv", 0) + active.get("wgrad", 0) > 0:
            tensor_busy += dt
    active[c] = active.get(c, 0) + d
    last = t
print("concurrency histogram (us with k kernels in flight):", {k: round(v, 1) for k, v in sorted(hist.items())})
print("busy us per class (any kernel of the class in flight):", {k: round(v, 1) for k, v in sorted(cls_busy.items())})
print("any tensor-core kernel in flight: %.1f us of %.1f" % (tensor_busy, s1 - s0))
sums = {}
for a, b, n in step:
    c = cls(n)
    sums[c] = sums.get(c, 0.0) + (b - a)
print("sum of kernel durations per class (us):", {k: round(v, 1) for k, v in sorted(sums.items())})
# per stream: busy time, number of kernels, and the longest idle gaps with the kernels on either side
per = {}
for a, b, n in step:
    per.setdefault(streams.get((a, n), -1), []).append((a, b, n))
for sid, ks in sorted(per.items(), key=lambda kv: -len(kv[1])):
    busy = sum(b - a for a, b, _ in ks)
    gaps = sorted(((ks[i + 1][0] - max(k[1] for k in ks[:i + 1][-3:]), ks[i][2][:40], ks[i + 1][2][:40], ks[i + 1][0] - s0)
                   for i in range(len(ks) - 1)), reverse=True)[:6]
    print("stream %s: %d kernels, busy %.0f us, first %.0f last %.0f" % (sid, len(ks), busy, ks[0][0] - s0, ks[-1][1] - s0))
    for g in gaps:
        if g[0] > 20:
            print("    gap %.0f us at t=%.0f  after %s  before %s" % (g[0], g[3], g[1], g[2]))
os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
json.dump({"step_us": s1 - s0, "hist": hist, "class_busy_us": cls_busy, "tensor_busy_us": tensor_busy, "class_sum_us": sums,
           "kernels": [(a - s0, b - s0, n[:60], streams.get((a, n), -1)) for a, b, n in step]}, open(args.out, "w"))
