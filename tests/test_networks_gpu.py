"""GPU parity of the six fused networks against the oracle (torch fp32 restatement of the reference,
run on the same device with TF32 disabled) on identical weights / inputs: outputs, input gradients
and every parameter gradient.

Tolerances.  Single layers meet north_star's 1e-3 bound (test_conv_gpu / test_wgrad_gpu / test_norm_gpu).
Through a whole network the reference's OWN reduced-precision paths are far from 1e-3 on gradients
(measured on B200, CIN generator at init, vs fp64: cuDNN TF32 dx 4.0e-2 / dW 5.8e-2; bf16 autocast
dx 0.19 / dW 0.32 -- InstanceNorm.scale ~ N(0,.02) makes activations tiny), so the bound is
stated the way SURVEY 9.3 recommends: our error against the fp32 oracle must be <= 2.5x the worst error of the
reference's own path at the same precision (cuDNN TF32 for tf32 mode, torch.autocast(bf16) for bf16
mode), computed in the same test on the same inputs; forward outputs additionally <= 2e-3 (tf32) /
3e-2 (bf16) rel-L2."""
import contextlib
import pytest
import torch

import dtg  # noqa: F401
from dtg_b200 import engine, networks
from oracle import nets as onets

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(a, b):
    return float((a.detach().float() - b.detach().float()).norm() / b.detach().float().norm().clamp_min(1e-12))


def _load(net, sd):
    missing, unexpected = net.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=False)
    assert not unexpected, unexpected
    for k in missing:   # only alias keys of the CIN res-blocks may be absent from the oracle dict
        assert k.split(".")[1] in ("10", "11", "12") and ".conv_block." not in k, k


def _oracle_params(sd):
    p = {}
    for k, v in sd.items():
        t = v.clone().to(DEV)
        if t.is_floating_point() and not ("running" in k):
            t.requires_grad_(True)
        p[k] = t
    return p


@pytest.fixture(autouse=True)
def _no_tf32():
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


STATE = onets.init_model_state(seed=1234, perturb=0.05)
OUT_TOL = {"tf32": 2e-3, "bf16": 3e-2}


@contextlib.contextmanager
def _lowprec(prec):
    """the reference's own reduced-precision path on this GPU"""
    if prec == "tf32":
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True
        try:
            yield
        finally:
            torch.backends.cudnn.allow_tf32 = False
            torch.backends.cuda.matmul.allow_tf32 = False
    else:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            yield


def _run_oracle(fn, sd, inputs, wgt, prec=None, head=0):
    """returns (out, [input grads], {param grads}) of the oracle; prec=None -> exact fp32"""
    op = _oracle_params(sd)
    ins = [t.detach().clone().requires_grad_(True) for t in inputs]
    with (_lowprec(prec) if prec else contextlib.nullcontext()):
        out = fn(op, *ins)
    outs = out if isinstance(out, tuple) else (out,)
    (outs[head].float() * wgt).sum().backward()
    return [o.detach().float() for o in outs], [t.grad for t in ins], {k: v.grad for k, v in op.items() if v.grad is not None}, op


def _compare(name, prec, net, ours_outs, ours_ingrads, ref, low):
    """ref / low: _run_oracle results at fp32 / reduced precision."""
    r_outs, r_in, r_pg, _ = ref
    l_outs, l_in, l_pg, _ = low
    for o, r in zip(ours_outs, r_outs):
        assert o.shape == r.shape
        assert _rel(o, r) < OUT_TOL[prec], (name, "out", _rel(o, r))
    real = [k for k in r_pg if not onets.is_noise_grad(name, k)]
    low_worst = max([_rel(l_pg[k], r_pg[k]) for k in real] + [_rel(a, b) for a, b in zip(l_in, r_in)])
    bound = 2.5 * low_worst + 2e-3
    for a, b in zip(ours_ingrads, r_in):
        assert _rel(a, b) < bound, (name, "dx", _rel(a, b), bound)
    pg = dict(net.named_parameters())
    for k in r_pg:
        if onets.is_noise_grad(name, k):
            # exact-zero gradient (bias in front of a mean-removing norm); the reference holds fp32 noise
            assert float(pg[k].grad.abs().max()) <= 1e-3 * (1.0 + float(r_pg[k].abs().max())), k
            continue
        assert _rel(pg[k].grad, r_pg[k]) < bound, (name, k, _rel(pg[k].grad, r_pg[k]), bound)


@pytest.mark.parametrize("prec", ["tf32", "bf16"])
def test_cin_generator(prec):
    engine.set_precision(prec)
    sd = STATE["netG_A_B"]
    net = networks.CINResnetGenerator(16, 3, 3, 32).to(DEV)
    _load(net, sd)
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(3, 3, 64, 64, generator=g) * 2 - 1).to(DEV).requires_grad_(True)
    z = torch.randn(3, 16, 1, 1, generator=g).to(DEV).requires_grad_(True)
    wgt = torch.randn(3, 3, 64, 64, generator=g).to(DEV)
    y = net(x, z)
    (y * wgt).sum().backward()
    ref = _run_oracle(onets.cin_resnet_generator, sd, [x, z], wgt)
    low = _run_oracle(onets.cin_resnet_generator, sd, [x, z], wgt, prec)
    _compare("netG_A_B", prec, net, [y], [x.grad, z.grad], ref, low)


@pytest.mark.parametrize("prec", ["tf32", "bf16"])
def test_resnet_generator(prec):
    engine.set_precision(prec)
    sd = STATE["netG_B_A"]
    net = networks.ResnetGenerator(3, 3, 32).to(DEV)
    _load(net, sd)
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(2, 3, 64, 64, generator=g) * 2 - 1).to(DEV).requires_grad_(True)
    wgt = torch.randn(2, 3, 64, 64, generator=g).to(DEV)
    y = net(x)
    (y * wgt).sum().backward()
    ref = _run_oracle(onets.resnet_generator, sd, [x], wgt)
    low = _run_oracle(onets.resnet_generator, sd, [x], wgt, prec)
    _compare("netG_B_A", prec, net, [y], [x.grad], ref, low)


@pytest.mark.parametrize("prec", ["tf32", "bf16"])
@pytest.mark.parametrize("which,size", [("netD_A", 64), ("netD_B", 64), ("netD_B", 67)])
def test_discriminators(which, prec, size):
    """size 67: odd extents take the ordinary strided first layer instead of the space-to-depth one"""
    engine.set_precision(prec)
    sd = STATE[which]
    net = (networks.Discriminator_edges(3, 32, norm_layer=networks.get_norm_layer("instance")) if which == "netD_A"
           else networks.Discriminator(3, 64, norm_layer=networks.get_norm_layer("instance"))).to(DEV)
    _load(net, sd)
    g = torch.Generator().manual_seed(2)
    x = (torch.rand(4, 3, size, size, generator=g) * 2 - 1).to(DEV).requires_grad_(True)
    y = net(x)
    wgt = torch.randn(y.shape, generator=g).to(DEV)
    (y * wgt).sum().backward()
    fn = onets.discriminator_edges if which == "netD_A" else onets.discriminator
    ref = _run_oracle(fn, sd, [x], wgt)
    low = _run_oracle(fn, sd, [x], wgt, prec)
    _compare(which, prec, net, [y], [x.grad], ref, low)


@pytest.mark.parametrize("prec", ["tf32", "bf16"])
def test_latent_nets(prec):
    engine.set_precision(prec)
    g = torch.Generator().manual_seed(3)
    # latent discriminator (Linear / BatchNorm1d)
    sd = STATE["netD_z_B"]
    net = networks.DiscriminatorLatent(16, 64).to(DEV)
    _load(net, sd)
    z = torch.randn(9, 16, 1, 1, generator=g).to(DEV).requires_grad_(True)
    y = net(z)
    wgt = torch.randn(9, 1, generator=g).to(DEV)
    (y * wgt).sum().backward()
    ref = _run_oracle(onets.discriminator_latent, sd, [z], wgt)
    low = _run_oracle(onets.discriminator_latent, sd, [z], wgt, prec)
    _compare("netD_z_B", prec, net, [y], [z.grad], ref, low)
    op = ref[3]
    assert _rel(net.model[1].running_mean, op["model.1.running_mean"]) < 2e-2
    assert _rel(net.model[7].running_var, op["model.7.running_var"]) < 2e-2
    assert int(net.model[1].num_batches_tracked) == 1
    # encoder (BatchNorm2d, two heads; only mu receives gradient like the default training step)
    sd = STATE["netE_B"]
    enc = networks.LatentEncoder(16, 6, 32, norm_layer=networks.get_norm_layer("batch")).to(DEV)
    _load(enc, sd)
    x = (torch.rand(5, 6, 64, 64, generator=g) * 2 - 1).to(DEV).requires_grad_(True)
    mu, lv = enc(x)
    wgt = torch.randn(5, 16, generator=g).to(DEV)
    (mu * wgt).sum().backward()
    ref = _run_oracle(onets.latent_encoder, sd, [x], wgt)
    low = _run_oracle(onets.latent_encoder, sd, [x], wgt, prec)
    assert mu.shape == (5, 16)
    _compare("netE_B", prec, enc, [mu, lv], [x.grad], ref, low)
    op = ref[3]
    assert _rel(enc.conv_modules[3].running_mean, op["conv_modules.3.running_mean"]) < 2e-2
    assert _rel(enc.conv_modules[12].running_var, op["conv_modules.12.running_var"]) < 2e-2


@pytest.mark.parametrize("prec", ["tf32", "bf16"])
def test_inference_sweep_config5(prec):
    """BASELINE config 5 (evaluate.py-style sweep, model.py:687-696 / 647-662): one input x 64 sampled z through
    G_A_B (generate_multi), plus predict_enc_params (E_B on cat(real_A, real_B)); forward-only, against the oracle."""
    import argparse
    from dtg_b200 import model as dmodel
    from oracle import step as ostep
    engine.set_precision(prec)
    opt = argparse.Namespace(**vars(ostep.default_opt()), expr_dir="/tmp", niter_decay=25)
    m = dmodel.AugmentedCycleGAN(opt, testing=True)
    for name, net in m._nets().items():
        _load(net, STATE[name])
    g = torch.Generator().manual_seed(5)
    real_A = (torch.rand(1, 3, 64, 64, generator=g) * 2 - 1).to(DEV)
    zs = torch.randn(64, 16, 1, 1, generator=g).to(DEV)
    out = m.generate_multi(real_A, zs)
    assert out.shape == (64, 3, 64, 64)
    with torch.no_grad():
        ref = onets.cin_resnet_generator(_oracle_params(STATE["netG_A_B"]), real_A.repeat(64, 1, 1, 1), zs)
    assert _rel(out, ref) < OUT_TOL[prec]
    a = (torch.rand(8, 3, 64, 64, generator=g) * 2 - 1).to(DEV)
    b = (torch.rand(8, 3, 64, 64, generator=g) * 2 - 1).to(DEV)
    (mu,) = m.predict_enc_params(a, b)
    with torch.no_grad():
        rmu, _ = onets.latent_encoder(_oracle_params(STATE["netE_B"]), torch.cat([a, b], 1))
    assert mu.shape == (8, 16) and _rel(mu, rmu.reshape(8, -1)) < OUT_TOL[prec]
    fake_A = m.predict_A(b)
    with torch.no_grad():
        rA = onets.resnet_generator(_oracle_params(STATE["netG_B_A"]), b)
    assert _rel(fake_A, rA) < OUT_TOL[prec]


@pytest.mark.parametrize("size", [128])
def test_fully_convolutional_nets_at_128(size):
    """BASELINE configs 3/4 shapes: generators and PatchGAN discriminators are fully convolutional;
    forward + backward at 128x128 (climate-field shaped: 3 -> 1 channels for the deterministic generator)."""
    prec = "bf16"
    engine.set_precision(prec)
    g = torch.Generator().manual_seed(9)
    sd = STATE["netG_A_B"]
    net = networks.CINResnetGenerator(16, 3, 3, 32).to(DEV)
    _load(net, sd)
    x = (torch.rand(2, 3, size, size, generator=g) * 2 - 1).to(DEV).requires_grad_(True)
    z = torch.randn(2, 16, 1, 1, generator=g).to(DEV).requires_grad_(True)
    wgt = torch.randn(2, 3, size, size, generator=g).to(DEV)
    y = net(x, z)
    (y * wgt).sum().backward()
    ref = _run_oracle(onets.cin_resnet_generator, sd, [x, z], wgt)
    low = _run_oracle(onets.cin_resnet_generator, sd, [x, z], wgt, prec)
    _compare("netG_A_B", prec, net, [y], [x.grad, z.grad], ref, low)
    sd = STATE["netD_B"]
    d = networks.Discriminator(3, 64, norm_layer=networks.get_norm_layer("instance")).to(DEV)
    _load(d, sd)
    xd = (torch.rand(2, 3, size, size, generator=g) * 2 - 1).to(DEV).requires_grad_(True)
    yd = d(xd)
    assert yd.shape == (2, 1, size // 4 - 3, size // 4 - 3)
    wd = torch.randn(*yd.shape, generator=g).to(DEV)
    (yd * wd).sum().backward()
    refd = _run_oracle(onets.discriminator, sd, [xd], wd)
    lowd = _run_oracle(onets.discriminator, sd, [xd], wd, prec)
    _compare("netD_B", prec, d, [yd], [xd.grad], refd, lowd)
