"""Per-op breakdown of one train_instance step: every ops.* call of one eager step is recorded with its
shape signature, then each UNIQUE call is replayed back to back (3 warm + R timed launches between two CUDA
events) so small kernels are not inflated by launch gaps.  Prints a table sorted by total time per step.

usage: python tools/step_profile.py [--batch 80] [--reps 10] [--json gpurun_out/step_profile.json]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import dtg  # noqa: E402,F401
from dtg_b200 import _lib as L, engine, model as dmodel, ops  # noqa: E402
from oracle import step as ostep  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=80)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--json", default=None)
args = ap.parse_args()

engine.set_precision(args.precision)
opt = argparse.Namespace(**vars(ostep.default_opt()), expr_dir="/tmp", niter_decay=25)
torch.manual_seed(1234)
m = dmodel.AugmentedCycleGAN(opt, testing=True)
m.prepare()
m.lanes.enabled = False      # serial issue: ops are replayed one by one below
a, b, z = [t.cuda() for t in ostep.synthetic_batch(args.batch, seed=4321)]
for _ in range(2):
    m._step_device(a, b, z)
torch.cuda.synchronize()


def sig(x):
    if isinstance(x, ops.PlaneT):
        return "P[%d,%d,%d,%d,h%d]" % (x.n, x.h, x.w, x.c, x.halo)
    if isinstance(x, torch.Tensor):
        return "T" + str(list(x.shape))
    if isinstance(x, (list, tuple)):
        return "[" + ",".join(sig(i) for i in x) + "]"
    if x is None:
        return "-"
    return str(x)


calls = []
NAMES = ["conv", "conv_wgrad", "norm_fwd", "norm_bwd", "cin_affine_fwd", "cin_affine_bwd", "pack_nchw", "unpack_nchw",
         "grad_gather", "channel_sum", "head1_fwd", "head1_dgrad", "head1_wgrad", "s2d_unfold_add", "loss_lsgan", "loss_l1", "loss_fused", "grad_sumsq", "adam_clip", "step_increment"]
orig = {n: getattr(ops, n) for n in NAMES}


def wrap(name):
    f = orig[name]

    def g(*a, **k):
        skip = {"ws", "scalars"}
        key = name + "(" + ",".join(sig(x) for x in a) + "," + ",".join("%s=%s" % (kk, sig(v)) for kk, v in sorted(k.items()) if kk not in skip) + ")"
        calls.append((key, f, a, k))
        return f(*a, **k)
    return g


for n in NAMES:
    setattr(ops, n, wrap(n))
orig_pack_run = ops.PackTable.run


def pack_run(self):
    calls.append(("pack_weights(%d items)" % len(self.items), orig_pack_run, (self,), {}))
    return orig_pack_run(self)


ops.PackTable.run = pack_run
snap = m._snapshot()
lc0 = L.lib().dtg_launch_count()
m._step_device(a, b, z)
nlaunch = L.lib().dtg_launch_count() - lc0
torch.cuda.synchronize()
for n in NAMES:
    setattr(ops, n, orig[n])
ops.PackTable.run = orig_pack_run

uniq = {}
for key, f, aa, kk in calls:
    u = uniq.setdefault(key, {"count": 0, "f": f, "a": aa, "k": kk})
    u["count"] += 1

R = args.reps
side = torch.cuda.Stream()
for key, u in uniq.items():
    f, aa, kk = u["f"], u["a"], u["k"]
    lc = L.lib().dtg_launch_count()
    for _ in range(3):
        f(*aa, **kk)
    u["launches"] = (L.lib().dtg_launch_count() - lc) // 3
    torch.cuda.synchronize()
    # R back-to-back launches captured in one CUDA graph: pure GPU time, no host launch gaps
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for _ in range(R):
            f(*aa, **kk)
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    u["us"] = e0.elapsed_time(e1) * 1e3 / R
    del g
m._restore(snap)


def flops(key, u):
    aa, kk = u["a"], u["k"]
    if key.startswith("conv("):
        x = aa[0]
        pix = x.n * (kk["out_h"] * kk["out_w"] if kk.get("mode", L.CONV_FWD) == L.CONV_FWD else x.h * x.w)
        return 2.0 * pix * (kk.get("cin") or aa[1].shape[2]) * kk["cout"] * kk["kh"] * kk["kw"]
    if key.startswith("conv_wgrad("):
        p = aa[0]
        return 2.0 * p.n * p.h * p.w * kk["pa"] * kk["qb"] * kk["kh"] * kk["kw"]
    return 0.0


rows = []
for key, u in uniq.items():
    fl = flops(key, u)
    rows.append({"key": key, "count": u["count"], "us": u["us"], "total_us": u["us"] * u["count"], "launches": u["launches"],
                 "tflops": fl / u["us"] / 1e6 if fl else None})
rows.sort(key=lambda r: -r["total_us"])
tot = sum(r["total_us"] for r in rows)
print("calls/step %d, unique %d, launches/step %d, sum of replayed times %.2f ms" % (len(calls), len(uniq), nlaunch, tot / 1e3))
by_fn = {}
for r in rows:
    fn = r["key"].split("(")[0]
    d = by_fn.setdefault(fn, [0, 0.0])
    d[0] += r["count"]
    d[1] += r["total_us"]
for fn, (c, t) in sorted(by_fn.items(), key=lambda kv: -kv[1][1]):
    print("  %-16s calls %4d  %9.1f us  %5.1f%%" % (fn, c, t, 100 * t / tot))
print()
for r in rows:
    print("%8.1f us x%3d = %9.1f us %5.1f%% %s %s" % (r["us"], r["count"], r["total_us"], 100 * r["total_us"] / tot,
                                                    ("%6.1f TF/s" % r["tflops"]) if r["tflops"] else "           ", r["key"]))
if args.json:
    json.dump({"batch": args.batch, "launches_per_step": int(nlaunch), "sum_ms": tot / 1e3, "rows": rows}, open(args.json, "w"), indent=1)
