"""Noise floor of the fused step under a mathematically neutral change: the same batch in a permuted sample order
(losses are batch means, gradients are sums over samples: both invariant).  What differs is only floating-point
summation order (split-K extents, batch-norm partial sums).  Used to interpret tools/dp_parity.py: a data-parallel run
cannot agree with the single-GPU run better than the single-GPU run agrees with itself under this permutation.
usage: python tools/perm_noise.py [--precision tf32|bf16] [--batch 8]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import dtg  # noqa: E402,F401
from dtg_b200 import engine, model as dmodel  # noqa: E402
from oracle import nets as onets, step as ostep  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="tf32")
ap.add_argument("--batch", type=int, default=8)
args = ap.parse_args()
engine.set_precision(args.precision)
state = onets.init_model_state(seed=1234, perturb=0.05)
a, b, z = [t.cuda() for t in ostep.synthetic_batch(args.batch, seed=4321)]


def run(a, b, z):
    opt = argparse.Namespace(**vars(ostep.default_opt()), expr_dir="/tmp", niter_decay=25)
    m = dmodel.AugmentedCycleGAN(opt, testing=True)
    for name, net in m._nets().items():
        net.load_state_dict({k: v.clone() for k, v in state[name].items()}, strict=False)
    m.prepare()
    for net in m._nets().values():
        net._ex.repack()
    m.train_instance(a, b, z)
    torch.cuda.synchronize()
    return {name: net._ex.arena.grad[:net._ex.arena.active_count].clone() for name, net in m._nets().items()}


g0 = run(a, b, z)
perm = torch.arange(args.batch - 1, -1, -1, device="cuda")           # reversed sample order
g1 = run(a[perm].contiguous(), b[perm].contiguous(), z[perm].contiguous())
g2 = run(a, b, z)                                                     # identical order again: must be bitwise equal
rel = lambda x, y: float((x - y).norm() / y.norm().clamp_min(1e-20))
print(json.dumps({"precision": args.precision, "batch": args.batch,
                  "permuted_vs_original_grad_rel": {n: rel(g1[n], g0[n]) for n in g0},
                  "repeat_bitwise_equal": all(bool(torch.equal(g2[n], g0[n])) for n in g0)}))
