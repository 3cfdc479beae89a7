"""Generate tests/golden/*.pt from the LIVE reference (/root/reference) -- run in the build container.

    python tests/golden/make_golden.py [aug|stoch]

The reference ships no golden vectors, so these are produced by executing the
unmodified reference modules (oracle.live_reference) on seeded weights / inputs that the tests can
regenerate without the reference (oracle.nets.init_model_state, oracle.step.synthetic_batch).
Stored: full small outputs for the network forwards, scalars + strided samples for the step.
"""
import copy
import os
import sys
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import live_reference as lr, nets, step  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SEED_W, SEED_X, PERTURB = 1234, 4321, 0.05


def sample(t, k=257):
    """Deterministic strided sample of a tensor (keeps fixtures small)."""
    f = t.detach().reshape(-1).float()
    return f[:: max(1, f.numel() // k)][:k].clone()


def main():
    warnings.simplefilter("ignore")
    torch.set_num_threads(8)
    opt = step.default_opt()
    state = nets.init_model_state(seed=SEED_W, perturb=PERTURB)
    ref = lr.build_reference_model(copy.deepcopy(opt), state)
    n = 2
    a, b, z = step.synthetic_batch(n, seed=SEED_X)

    # ---- per-network forward / input-gradient fixtures ---------------------------------
    fwd = {}
    with torch.enable_grad():
        ar = a.clone().requires_grad_(True); zr = z.clone().requires_grad_(True)
        y = ref.netG_A_B(ar, zr); (y * torch.linspace(-1, 1, y.numel()).view_as(y)).sum().backward()
        fwd["G_A_B"] = dict(out=y.detach().clone(), dx=sample(ar.grad), dz=zr.grad.clone(),
                            dw_norm={k: float(p.grad.norm()) for k, p in ref.netG_A_B.named_parameters()})
        br = b.clone().requires_grad_(True)
        y = ref.netG_B_A(br); (y * torch.linspace(-1, 1, y.numel()).view_as(y)).sum().backward()
        fwd["G_B_A"] = dict(out=y.detach().clone(), dx=sample(br.grad),
                            dw_norm={k: float(p.grad.norm()) for k, p in ref.netG_B_A.named_parameters()})
        for nm, net, x in (("D_A", ref.netD_A, a), ("D_B", ref.netD_B, b)):
            xr = x.clone().requires_grad_(True)
            y = net(xr); (y * torch.linspace(-1, 1, y.numel()).view_as(y)).sum().backward()
            fwd[nm] = dict(out=y.detach().clone(), dx=sample(xr.grad),
                           dw_norm={k: float(p.grad.norm()) for k, p in net.named_parameters()})
    for net in (ref.netG_A_B, ref.netG_B_A, ref.netD_A, ref.netD_B):
        net.zero_grad()
    # encoder / latent discriminator are BatchNorm nets: reload state afterwards (running stats move)
    ref = lr.build_reference_model(copy.deepcopy(opt), state)
    enc_in = torch.cat((a, b), 1)
    mu, lv = ref.netE_B(enc_in)
    fwd["E_B"] = dict(mu=mu.detach().clone(), logvar=lv.detach().clone(),
                      running_mean_3=ref.netE_B.state_dict()["conv_modules.3.running_mean"].clone(),
                      running_var_12=ref.netE_B.state_dict()["conv_modules.12.running_var"].clone())
    fwd["D_z_B"] = dict(out=ref.netD_z_B(z).detach().clone())
    torch.save(dict(n=n, seed_w=SEED_W, seed_x=SEED_X, perturb=PERTURB, nets=fwd),
               os.path.join(HERE, "golden_nets_n2.pt"))

    # ---- train_instance fixtures (2 consecutive steps) ----------------------------------
    ref = lr.build_reference_model(copy.deepcopy(opt), state)
    steps = []
    for it in range(2):
        losses, visuals, gnorms = ref.train_instance(a, b, z)
        rec = dict(losses={k: float(v) for k, v in losses.items()},
                   gnorms={k: float(v) for k, v in gnorms.items()},
                   visuals={k: sample(v, 1025) for k, v in visuals.items()},
                   grad_norm={}, param_norm={})
        for name in nets.NET_NAMES:
            net = getattr(ref, name)
            for k, p in net.named_parameters():
                if ".conv_block." not in k and k.split(".")[1] in ("10", "11", "12") and name.startswith("netG"):
                    continue  # alias of a conv_block.* key
                rec["param_norm"]["%s/%s" % (name, k)] = float(p.detach().double().norm())
                if name in ("netG_A_B", "netG_B_A", "netE_B") and p.grad is not None:
                    rec["grad_norm"]["%s/%s" % (name, k)] = float(p.grad.double().norm())
        steps.append(rec)
    torch.save(dict(n=n, seed_w=SEED_W, seed_x=SEED_X, perturb=PERTURB, steps=steps),
               os.path.join(HERE, "golden_step_n2.pt"))
    for f in ("golden_nets_n2.pt", "golden_step_n2.pt"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


def main_stoch():
    """StochCycleGAN.train_instance (model.py:126-208) on BASELINE config 3's shape (climate fields, 3 -> 1
    channels, 128x128) and AugmentedCycleGAN.supervised_train_instance (model.py:541-604) at 64x64."""
    warnings.simplefilter("ignore")
    torch.set_num_threads(8)
    n, size = 2, 128
    opt = step.default_opt(output_nc=1)
    state = nets.init_model_state(seed=SEED_W, perturb=PERTURB, output_nc=1)
    ref = lr.build_reference_model(copy.deepcopy(opt), state, stoch=True)
    a, b, z = step.synthetic_batch(n, size=size, seed=SEED_X, output_nc=1, kind="climate")
    steps = []
    for it in range(2):
        losses, visuals, gnorms = ref.train_instance(a, b, z)
        rec = dict(losses={k: float(v) for k, v in losses.items()},
                   gnorms={k: float(v) for k, v in gnorms.items()},
                   visuals={k: sample(v, 1025) for k, v in visuals.items()}, param_norm={})
        for name in ("netG_A_B", "netG_B_A", "netD_A", "netD_B"):
            for k, p in getattr(ref, name).named_parameters():
                rec["param_norm"]["%s/%s" % (name, k)] = float(p.detach().double().norm())
        steps.append(rec)
    torch.save(dict(n=n, size=size, output_nc=1, kind="climate", seed_w=SEED_W, seed_x=SEED_X, perturb=PERTURB,
                    steps=steps), os.path.join(HERE, "golden_stoch_n2.pt"))

    opt = step.default_opt()
    state = nets.init_model_state(seed=SEED_W, perturb=PERTURB)
    ref = lr.build_reference_model(copy.deepcopy(opt), state)
    a, b, z = step.synthetic_batch(n, seed=SEED_X)
    sup = []
    for it in range(2):
        losses = ref.supervised_train_instance(a, b, z)
        sup.append({k: float(v) for k, v in losses.items()})
    torch.save(dict(n=n, seed_w=SEED_W, seed_x=SEED_X, perturb=PERTURB, steps=sup),
               os.path.join(HERE, "golden_sup_n2.pt"))
    for f in ("golden_stoch_n2.pt", "golden_sup_n2.pt"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    if len(sys.argv) < 2 or sys.argv[1] == "aug":
        main()
    if len(sys.argv) < 2 or sys.argv[1] == "stoch":
        main_stoch()
