import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dtg  # noqa
from dtg_b200 import engine, evaluate as ev, model as dmodel
from oracle import nets as onets, step as ostep
engine.set_precision("tf32")
opt = argparse.Namespace(**vars(ostep.default_opt()), expr_dir="/tmp", niter_decay=25)
state = onets.init_model_state(seed=5, perturb=0.03)
m = dmodel.AugmentedCycleGAN(opt, testing=True)
for name, net in m._nets().items():
    net.load_state_dict({k: v.clone() for k, v in state[name].items()}, strict=False)
m.prepare()
for net in m._nets().values():
    net._ex.repack()
a, b, _ = [t.cuda() for t in ostep.synthetic_batch(4, seed=9)]
for steps in (1, 2, 3):
    qs = []
    for kw in ({}, {"compute_l1": True}):
        torch.manual_seed(3)
        q = {}
        r = ev.variational_ubo(m, a, b, steps, q_out=q, **kw)
        qs.append((r, q))
    (r0, q0), (r1, q1) = qs
    print(steps, r0, r1)
    for k in ("mu", "logvar"):
        d = (q0[k] - q1[k]).abs()
        print("  ", k, "max diff", float(d.max()), "mean", float(d.mean()), "n>1e-3:", int((d > 1e-3).sum()), "of", d.numel())

om = ostep.OracleModel(ostep.default_opt(), state, device="cuda")
class Adapter(object):
    opt = om.opt
    netE_B = True
    def predict_B(self, x, z):
        return om.G_A_B(x, z)
    def predict_enc_params(self, x, y):
        with torch.no_grad():
            mu, _ = om.E_B(torch.cat((x, y), 1))
        return (mu.reshape(mu.shape[0], -1),)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
for steps in (1,):
    res = {}
    for name, mod, kw in (("fused", m, {}), ("generic", m, {"compute_l1": True}), ("oracle", Adapter(), {})):
        torch.manual_seed(3)
        q = {}
        ev.variational_ubo(mod, a, b, steps, q_out=q, **kw)
        res[name] = q
    for x, y in (("fused", "oracle"), ("generic", "oracle"), ("fused", "generic")):
        for k in ("mu", "logvar"):
            d = (res[x][k] - res[y][k]).abs()
            print(x, "vs", y, k, "max", float(d.max()), "n>1e-3:", int((d > 1e-3).sum()))
# direct dz comparison
z = torch.randn(4, 16, 1, 1, device="cuda").requires_grad_(True)
w = torch.randn(4, 3, 64, 64, device="cuda")
y = m.predict_B(a, z); (y * w).sum().backward(); dz_g = z.grad.clone().view(4, 16)
z2 = z.detach().clone().requires_grad_(True)
y2 = om.G_A_B(a, z2); (y2 * w).sum().backward(); dz_o = z2.grad.view(4, 16)
print("dz generic vs oracle rel", float((dz_g - dz_o).norm() / dz_o.norm()), "abs dz", float(dz_o.abs().mean()))
