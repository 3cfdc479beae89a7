"""Permutation noise of the ORACLE (cuDNN / ATen, fp32 and TF32) next to ours, and ours-vs-oracle at the same batch:
is a large permutation noise of the BatchNorm-coupled networks a property of the problem or of our kernels?
usage: python tools/perm_noise_oracle.py [--batch 8]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dtg  # noqa
from dtg_b200 import engine, model as dmodel
from oracle import nets as onets, step as ostep

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--seed", type=int, default=4321)
ap.add_argument("--kind", default="edges2shoes")
ap.add_argument("--wseed", type=int, default=1234)
ap.add_argument("--wscale", type=float, default=1.0)
args = ap.parse_args()
state = onets.init_model_state(seed=args.wseed, perturb=0.05)
if args.wscale != 1.0:      # larger conv weights: features vary more across samples (BatchNorm better conditioned)
    for sd in state.values():
        for k, v in sd.items():
            if k.endswith(".weight") and v.dim() == 4 and "conv" not in k.split(".")[-2]:
                v.mul_(args.wscale)
a, b, z = [t.cuda() for t in ostep.synthetic_batch(args.batch, seed=args.seed, kind=args.kind)]
perm = torch.arange(args.batch - 1, -1, -1, device="cuda")
rel = lambda x, y: float((x - y).norm() / y.norm().clamp_min(1e-20))


def oracle(a, b, z, tf32):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    om = ostep.OracleModel(ostep.default_opt(), state, device="cuda")
    grabbed = {}
    def grab(nets):
        def f(m):
            for net in nets:
                grabbed[net] = torch.cat([v.grad.reshape(-1) for k, v in m.params(net) if v.grad is not None])
        return f
    om.train_instance(a, b, z, hooks={"after_D_backward": grab(("netD_A", "netD_B", "netD_z_B")),
                                      "after_G_backward": grab(("netG_A_B", "netG_B_A", "netE_B"))})
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return grabbed


def ours(a, b, z, prec):
    engine.set_precision(prec)
    opt = argparse.Namespace(**vars(ostep.default_opt()), expr_dir="/tmp", niter_decay=25)
    m = dmodel.AugmentedCycleGAN(opt, testing=True)
    for name, net in m._nets().items():
        net.load_state_dict({k: v.clone() for k, v in state[name].items()}, strict=False)
    m.prepare()
    for net in m._nets().values():
        net._ex.repack()
    m.train_instance(a, b, z)
    torch.cuda.synchronize()
    out = {}
    for name, net in m._nets().items():
        out[name] = torch.cat([p.grad.reshape(-1) for k, p in net.named_parameters() if p.grad is not None])
    return out


res = {}
for tag, tf32 in (("oracle_fp32", False), ("oracle_tf32", True)):
    g0, g1 = oracle(a, b, z, tf32), oracle(a[perm].contiguous(), b[perm].contiguous(), z[perm].contiguous(), tf32)
    res[tag + "_perm_noise"] = {n: round(rel(g1[n], g0[n]), 6) for n in g0}
    res[tag] = g0
for prec in ("tf32", "bf16"):
    g0, g1 = ours(a, b, z, prec), ours(a[perm].contiguous(), b[perm].contiguous(), z[perm].contiguous(), prec)
    res["ours_%s_perm_noise" % prec] = {n: round(rel(g1[n], g0[n]), 6) for n in g0}
    res["ours_%s_vs_oracle_fp32" % prec] = {n: round(rel(g0[n], res["oracle_fp32"][n]), 6) for n in g0 if g0[n].numel() == res["oracle_fp32"][n].numel()}
res["oracle_tf32_vs_oracle_fp32"] = {n: round(rel(res["oracle_tf32"][n], res["oracle_fp32"][n]), 6) for n in res["oracle_fp32"]}
del res["oracle_fp32"], res["oracle_tf32"]
for k, v in res.items():
    print(k, json.dumps(v))
