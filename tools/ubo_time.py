"""variational_ubo (evaluate.py:39-148) on the fused model: fused loop vs the PyTorch objective around predict_B.
usage: python tools/ubo_time.py [batch=80] [steps=50]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dtg  # noqa
from dtg_b200 import engine, evaluate as ev, model as dmodel
from oracle import step as ostep

n = int(sys.argv[1]) if len(sys.argv) > 1 else 80
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
engine.set_precision("bf16")
opt = argparse.Namespace(**vars(ostep.default_opt()), expr_dir="/tmp", niter_decay=25)
torch.manual_seed(0)
m = dmodel.AugmentedCycleGAN(opt, testing=True)
m.prepare()
a, b, _ = [t.cuda() for t in ostep.synthetic_batch(n, seed=9)]
for name, kw in (("fused", {}), ("pytorch-objective", {"compute_l1": True})):
    torch.manual_seed(1)
    ev.variational_ubo(m, a, b, 3, **kw)
    torch.cuda.synchronize()
    torch.manual_seed(1)
    t0 = time.perf_counter()
    r = ev.variational_ubo(m, a, b, steps, **kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("variational_ubo %-18s batch %d, %d steps: %.1f ms/step  (ubo %.2f kld %.3f bpp %.4f)" % (name, n, steps, dt * 1e3 / steps, *r))
