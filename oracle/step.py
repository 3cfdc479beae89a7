"""Oracle restatement of ``AugmentedCycleGAN.train_instance`` (model.py:402-539),
``AugmentedCycleGAN.supervised_train_instance`` (model.py:541-604) and
``StochCycleGAN.train_instance`` (model.py:126-208).

Test infrastructure only.  ``OracleModel`` holds the six parameter dicts (reference
state_dict keys), per-parameter Adam state, and executes the step with torch autograd on
CPU (or any device), mirroring the reference order of operations exactly: 15 network
forwards, D backward / clip / Adam, G backward / clip / Adam, reporting dicts.
"""
from collections import OrderedDict
from types import SimpleNamespace

import torch
import torch.nn.functional as F

from . import nets
from .functional import gauss_reparametrize, kld_std_gauss, log_prob_gaussian, lsgan


def default_opt(**kw):
    """Hot-path defaults of options.py:22-83."""
    o = dict(input_nc=3, output_nc=3, ngf=32, nef=32, ndf=64, nlatent=16, lr=2e-4, beta1=0.5,
             max_gnorm=500.0, stoch_enc=False, z_gan=1, enc_A_B=1, lambda_A=1.0, lambda_B=1.0,
             lambda_z_B=0.025, lambda_sup_A=0.1, lambda_sup_B=0.1, monitor_gnorm=True, no_lsgan=False, use_dropout=False,
             norm="instance", which_model_netG="resnet", which_model_netD="basic", gpu_ids=[])
    o.update(kw)
    return SimpleNamespace(**o)


def _is_param(k):
    return not (k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked"))


class OracleModel:
    # optimizer grouping and learning rates: model.py:379-389
    GROUPS = OrderedDict([("G_A", ("netG_B_A",)), ("G_B", ("netG_A_B", "netE_B")),
                          ("D_A", ("netD_A",)), ("D_B", ("netD_B", "netD_z_B"))])

    def __init__(self, opt=None, state=None, device="cpu", dtype=torch.float32):
        self.opt = opt or default_opt()
        o = self.opt
        state = state or nets.init_model_state(input_nc=o.input_nc, output_nc=o.output_nc, ngf=o.ngf,
                                               nef=o.nef, ndf=o.ndf, nlatent=o.nlatent, enc_A_B=bool(o.enc_A_B))
        self.nets = {}
        for name in nets.NET_NAMES:
            sd = {}
            for k, v in state[name].items():
                t = v.detach().clone().to(device)
                if t.is_floating_point():
                    t = t.to(dtype)
                    if _is_param(k):
                        t.requires_grad_(True)
                sd[k] = t
            self.nets[name] = sd
        self.lr = {"G_A": o.lr, "G_B": o.lr, "D_A": o.lr / 5.0, "D_B": o.lr / 5.0}
        self.adam = {}  # (net, key) -> dict(step, m, v)

    # ---- network forwards -------------------------------------------------------------
    def G_A_B(self, a, z): return nets.cin_resnet_generator(self.nets["netG_A_B"], a, z)
    def G_B_A(self, b): return nets.resnet_generator(self.nets["netG_B_A"], b)
    def E_B(self, x): return nets.latent_encoder(self.nets["netE_B"], x)
    def D_A(self, x): return nets.discriminator_edges(self.nets["netD_A"], x)
    def D_B(self, x): return nets.discriminator(self.nets["netD_B"], x)
    def D_z_B(self, z): return nets.discriminator_latent(self.nets["netD_z_B"], z)

    def params(self, net):
        return [(k, v) for k, v in self.nets[net].items() if v.is_floating_point() and v.requires_grad]

    def _zero_grad(self, group):
        for net in self.GROUPS[group]:
            for _, v in self.params(net):
                v.grad = None

    def _clip(self, net):
        """torch.nn.utils.clip_grad_norm (model.py:447-449,510-512): params with grad only."""
        gs = [v.grad for _, v in self.params(net) if v.grad is not None]
        total = torch.sqrt(sum((g.double() ** 2).sum() for g in gs)).to(gs[0].dtype)
        coef = float(self.opt.max_gnorm) / (float(total) + 1e-6)
        if coef < 1.0:
            for g in gs:
                g.mul_(coef)
        return float(total)

    def _adam(self, group):
        """torch.optim.Adam.step, betas=(beta1, 0.999), eps=1e-8, no weight decay."""
        b1, b2, eps, lr = self.opt.beta1, 0.999, 1e-8, self.lr[group]
        with torch.no_grad():
            for net in self.GROUPS[group]:
                for k, v in self.params(net):
                    if v.grad is None:          # e.g. enc_logvar
                        continue
                    st = self.adam.setdefault((net, k), dict(step=0, m=torch.zeros_like(v), v=torch.zeros_like(v)))
                    st["step"] += 1
                    g = v.grad
                    st["m"].mul_(b1).add_(g, alpha=1 - b1)
                    st["v"].mul_(b2).addcmul_(g, g, value=1 - b2)
                    bc1 = 1 - b1 ** st["step"]
                    bc2 = 1 - b2 ** st["step"]
                    denom = (st["v"].sqrt() / (bc2 ** 0.5)).add_(eps)
                    v.addcdiv_(st["m"], denom, value=-lr / bc1)

    # ---- the step ----------------------------------------------------------------------
    def train_instance(self, real_A, real_B, prior_z_B, hooks=None):
        o = self.opt
        hooks = hooks or {}
        fake_B = self.G_A_B(real_A, prior_z_B)                                    # model.py:404
        fake_A = self.G_B_A(real_B)                                               # :407
        enc_in = torch.cat((fake_A, real_B), 1) if o.enc_A_B else real_B          # :409-413
        mu_z_realB, logvar_z_realB = self.E_B(enc_in)
        if o.stoch_enc:
            post_z_realB = gauss_reparametrize(mu_z_realB, logvar_z_realB)
        else:
            post_z_realB = mu_z_realB.reshape(mu_z_realB.shape[0], mu_z_realB.shape[1], 1, 1)
            logvar_z_realB = logvar_z_realB * 0.0

        def discriminate(net, fake, real):                                        # :327-334
            pf = net(fake); lf = lsgan(pf, False)
            pt = net(real); lt = lsgan(pt, True)
            return lf, lt, pf, pt

        lfA, ltA, pred_fake_A, pred_true_A = discriminate(self.D_A, fake_A.detach(), real_A)
        lfB, ltB, pred_fake_B, pred_true_B = discriminate(self.D_B, fake_B.detach(), real_B)
        lpz, lqz, _, _ = discriminate(self.D_z_B, post_z_realB.detach(), prior_z_B)
        loss_D_A = 0.5 * (lfA + ltA)
        loss_D_B = 0.5 * (lfB + ltB)
        loss_D_z_B = 0.5 * (lpz + lqz)
        loss_D = loss_D_A + loss_D_B
        if o.z_gan and not o.stoch_enc:
            loss_D = loss_D + loss_D_z_B
        self._zero_grad("D_A"); self._zero_grad("D_B")
        loss_D.backward()
        gnorm_D_A = self._clip("netD_A"); gnorm_D_B = self._clip("netD_B"); gnorm_D_z_B = self._clip("netD_z_B")
        if "after_D_backward" in hooks:
            hooks["after_D_backward"](self)
        self._adam("D_A"); self._adam("D_B")

        pred_fake_A = self.D_A(fake_A); loss_G_A = lsgan(pred_fake_A, True)       # :457-464
        pred_fake_B = self.D_B(fake_B); loss_G_B = lsgan(pred_fake_B, True)
        pred_post_z_B = self.D_z_B(post_z_realB); loss_G_z_B = lsgan(pred_post_z_B, True)
        rec_A = self.G_B_A(fake_B); loss_cycle_A = F.l1_loss(rec_A, real_A)       # :467-468
        enc_in2 = torch.cat((real_A, fake_B), 1) if o.enc_A_B else fake_B         # :471-475
        mu_z_fakeB, logvar_z_fakeB = self.E_B(enc_in2)
        bs = prior_z_B.shape[0]
        if o.stoch_enc:
            lp = log_prob_gaussian(prior_z_B.reshape(bs, o.nlatent), mu_z_fakeB.reshape(bs, o.nlatent),
                                   logvar_z_fakeB.reshape(bs, o.nlatent))
            loss_cycle_z_B = -1.0 * lp.mean(1).mean(0)
        else:
            loss_cycle_z_B = F.l1_loss(mu_z_fakeB.reshape(bs, o.nlatent), prior_z_B.reshape(bs, o.nlatent))
        kld_z_B = kld_std_gauss(mu_z_realB, logvar_z_realB).mean(0)               # :490
        rec_B = self.G_A_B(fake_A, post_z_realB); loss_cycle_B = F.l1_loss(rec_B, real_B)   # :493-494
        loss_cycle = loss_cycle_A * o.lambda_A + loss_cycle_B * o.lambda_B + loss_cycle_z_B * o.lambda_z_B
        loss_G = loss_G_A + loss_G_B + loss_cycle
        if o.stoch_enc:
            loss_G = loss_G + kld_z_B * o.lambda_z_B
        if o.z_gan and not o.stoch_enc:
            loss_G = loss_G + loss_G_z_B
        self._zero_grad("G_A"); self._zero_grad("G_B")
        loss_G.backward()
        gn_GAB = self._clip("netG_A_B"); gn_GBA = self._clip("netG_B_A"); gn_E = self._clip("netE_B")
        if "after_G_backward" in hooks:
            hooks["after_G_backward"](self)
        self._adam("G_A"); self._adam("G_B")

        f = lambda t: float(t.detach())
        losses = OrderedDict([("D_A", f(loss_D_A)), ("G_A", f(loss_G_A)), ("Cyc_A", f(loss_cycle_A)),
                              ("Cyc_z_B", f(loss_cycle_z_B)), ("KLD_z_B", f(kld_z_B)),
                              ("D_B", f(loss_D_B)), ("G_B", f(loss_G_B)), ("Cyc_B", f(loss_cycle_B)),
                              ("D_z_B", f(loss_D_z_B)),
                              ("P_t_A", f(pred_true_A.mean())), ("P_f_A", f(pred_fake_A.mean())),
                              ("P_t_B", f(pred_true_B.mean())), ("P_f_B", f(pred_fake_B.mean()))])
        visuals = OrderedDict([("real_A", real_A.detach()), ("fake_B", fake_B.detach()), ("rec_A", rec_A.detach()),
                               ("real_B", real_B.detach()), ("fake_A", fake_A.detach()), ("rec_B", rec_B.detach())])
        gnorms = OrderedDict([("gnorm_G_A_B", gn_GAB), ("gnorm_G_B_A", gn_GBA), ("gnorm_E_B", gn_E),
                              ("gnorm_D_B", gnorm_D_B), ("gnorm_D_z_B", gnorm_D_z_B), ("gnorm_D_A", gnorm_D_A),
                              ("mu_min", f(mu_z_realB.min())), ("mu_max", f(mu_z_realB.max())),
                              ("logvar_min", f(logvar_z_realB.min())), ("logvar_max", f(logvar_z_realB.max()))])
        if o.monitor_gnorm:
            return losses, visuals, gnorms
        return losses, visuals

    # ---- the supervised step (model.py:541-604) -----------------------------------------
    def supervised_train_instance(self, real_A, real_B, prior_z_B):
        o = self.opt
        enc_in = torch.cat((real_A, real_B), 1) if o.enc_A_B else real_B          # :543-547
        mu, logvar = self.E_B(enc_in)
        if o.stoch_enc:
            post_z_B = gauss_reparametrize(mu, logvar)
        else:
            post_z_B = mu.reshape(mu.shape[0], mu.shape[1], 1, 1)                 # :552
            logvar = logvar * 0.0
        lpz = lsgan(self.D_z_B(post_z_B.detach()), False)                         # :555-557
        lqz = lsgan(self.D_z_B(prior_z_B), True)
        loss_D_z_B = 0.5 * (lpz + lqz)
        self._zero_grad("D_B")                                                    # :559-562
        loss_D_z_B.backward()
        gnorm_D_z_B = self._clip("netD_z_B")
        self._adam("D_B")             # netD_B has no gradient here: torch skips its parameters
        pred_B = self.G_A_B(real_A, post_z_B)                                     # :564-568
        pred_A = self.G_B_A(real_B)
        loss_sup_A = F.l1_loss(pred_A, real_A)
        loss_sup_B = F.l1_loss(pred_B, real_B)
        loss_G_z_B = lsgan(self.D_z_B(post_z_B), True)                            # :570-571
        kld_z_B = kld_std_gauss(mu, logvar).mean(0)                               # :574
        loss_G = loss_sup_A * o.lambda_sup_A + loss_sup_B * o.lambda_sup_B        # :577
        if o.stoch_enc:
            loss_G = loss_G + kld_z_B * o.lambda_z_B
        if o.z_gan and not o.stoch_enc:
            loss_G = loss_G + loss_G_z_B
        self._zero_grad("G_A"); self._zero_grad("G_B")
        loss_G.backward()
        gn_GAB = self._clip("netG_A_B"); gn_GBA = self._clip("netG_B_A"); gn_E = self._clip("netE_B")
        self._adam("G_A"); self._adam("G_B")
        f = lambda t: float(t.detach())
        return OrderedDict([("S_A", f(loss_sup_A)), ("S_B", f(loss_sup_B)), ("KLD_z_B", f(kld_z_B)),
                            ("D_z_B", f(loss_D_z_B)), ("gnorm_G_A_B", gn_GAB), ("gnorm_G_B_A", gn_GBA),
                            ("gnorm_E_B", gn_E), ("gnorm_D_z_B", gnorm_D_z_B)])


class OracleStochModel(OracleModel):
    """StochCycleGAN (model.py:75-325): the two generators and two image discriminators only; one Adam
    for both generators (lr) and one for both discriminators (lr/5) (model.py:109-114).  Fully
    convolutional, so it is the reference-native step at 128x128 / 256x256."""
    GROUPS = OrderedDict([("G", ("netG_A_B", "netG_B_A")), ("D", ("netD_A", "netD_B"))])
    NETS = ("netG_A_B", "netG_B_A", "netD_A", "netD_B")

    def __init__(self, opt=None, state=None, device="cpu", dtype=torch.float32, ignore_noise=False):
        super().__init__(opt, state, device, dtype)
        for k in ("netE_B", "netD_z_B"):
            self.nets.pop(k)
        self.lr = {"G": self.opt.lr, "D": self.opt.lr / 5.0}
        self.ignore_noise = ignore_noise

    def train_instance(self, real_A, real_B, prior_z_B, hooks=None):
        o = self.opt
        hooks = hooks or {}
        if self.ignore_noise:
            prior_z_B = prior_z_B * 0.0 + 1.0                                     # model.py:128-129
        fake_B = self.G_A_B(real_A, prior_z_B)                                    # :132
        fake_A = self.G_B_A(real_B)                                               # :135
        pred_fake_A = self.D_A(fake_A.detach()); lfA = lsgan(pred_fake_A, False)  # :139-152
        pred_true_A = self.D_A(real_A); ltA = lsgan(pred_true_A, True)
        pred_fake_B = self.D_B(fake_B.detach()); lfB = lsgan(pred_fake_B, False)
        pred_true_B = self.D_B(real_B); ltB = lsgan(pred_true_B, True)
        loss_D_A = 0.5 * (lfA + ltA)
        loss_D_B = 0.5 * (lfB + ltB)
        loss_D = loss_D_A + loss_D_B
        self._zero_grad("D")
        loss_D.backward()
        gnorm_D_A = self._clip("netD_A"); gnorm_D_B = self._clip("netD_B")
        if "after_D_backward" in hooks:
            hooks["after_D_backward"](self)
        self._adam("D")
        pred_fake_A = self.D_A(fake_A); loss_G_A = lsgan(pred_fake_A, True)       # :168-172
        pred_fake_B = self.D_B(fake_B); loss_G_B = lsgan(pred_fake_B, True)
        rec_A = self.G_B_A(fake_B); loss_cycle_A = F.l1_loss(rec_A, real_A)       # :175-176
        rec_B = self.G_A_B(fake_A, prior_z_B); loss_cycle_B = F.l1_loss(rec_B, real_B)   # :179-180
        loss_cycle = loss_cycle_A * o.lambda_A + loss_cycle_B * o.lambda_B
        loss_G = loss_G_A + loss_G_B + loss_cycle
        self._zero_grad("G")
        loss_G.backward()
        gn_GAB = self._clip("netG_A_B"); gn_GBA = self._clip("netG_B_A")
        if "after_G_backward" in hooks:
            hooks["after_G_backward"](self)
        self._adam("G")
        f = lambda t: float(t.detach())
        losses = OrderedDict([("D_A", f(loss_D_A)), ("G_A", f(loss_G_A)), ("Cyc_A", f(loss_cycle_A)),
                              ("D_B", f(loss_D_B)), ("G_B", f(loss_G_B)), ("Cyc_B", f(loss_cycle_B)),
                              ("P_t_A", f(pred_true_A.mean())), ("P_f_A", f(pred_fake_A.mean())),
                              ("P_t_B", f(pred_true_B.mean())), ("P_f_B", f(pred_fake_B.mean()))])
        visuals = OrderedDict([("real_A", real_A.detach()), ("fake_B", fake_B.detach()), ("rec_A", rec_A.detach()),
                               ("real_B", real_B.detach()), ("fake_A", fake_A.detach()), ("rec_B", rec_B.detach())])
        gnorms = OrderedDict([("gnorm_G_A_B", gn_GAB), ("gnorm_G_B_A", gn_GBA),
                              ("gnorm_D_B", gnorm_D_B), ("gnorm_D_A", gnorm_D_A)])
        if o.monitor_gnorm:
            return losses, visuals, gnorms
        return losses, visuals

    def supervised_train_instance(self, *a, **k):
        raise AttributeError("StochCycleGAN has no supervised_train_instance (model.py:75-325)")


def synthetic_batch(n, size=64, seed=4321, input_nc=3, output_nc=3, nlatent=16, kind="edges2shoes"):
    """Synthetic edges2shoes-shaped data: A = sparse +-1 edge maps, B = smooth colour
    fields, both float32 NCHW in [-1,1]; z ~ N(0,1) [N,Z,1,1].
    kind="climate": Livneh-style gridded fields (BASELINE config 3): low-pass Gaussian random fields times a
    fixed land mask (~45 % ocean), scaled per sample and channel to [-1,1] with masked cells -> 0, the way
    dataloader.py:17-33 treats the NaN ocean cells of the .npz grids."""
    g = torch.Generator().manual_seed(seed)
    if kind == "climate":
        gm = torch.Generator().manual_seed(2718)                  # the mask is a property of the domain, not the batch
        k1 = torch.exp(-0.5 * ((torch.arange(33) - 16.0) / 8.0) ** 2)
        k1 = (k1 / k1.sum()).view(1, 1, 1, 33)

        def smooth(t):                                            # separable 8-px Gaussian blur, reflect borders
            c = t.shape[1]
            t = F.conv2d(F.pad(t, (16, 16, 0, 0), mode="reflect"), k1.expand(c, 1, 1, 33), groups=c)
            return F.conv2d(F.pad(t, (0, 0, 16, 16), mode="reflect"), k1.transpose(2, 3).expand(c, 1, 33, 1), groups=c)

        lf = smooth(torch.randn(1, 1, size, size, generator=gm))
        land = (lf > lf.flatten().kthvalue(int(0.45 * size * size)).values).float()

        def fields(c):
            f = smooth(torch.randn(n, c, size, size, generator=g))
            lo = f.amin(dim=(2, 3), keepdim=True); hi = f.amax(dim=(2, 3), keepdim=True)
            return ((f - lo) / (hi - lo).clamp_min(1e-12) * 2 - 1) * land

        a, b = fields(input_nc), fields(output_nc)
    elif kind == "uniform":
        a = torch.rand(n, input_nc, size, size, generator=g) * 2 - 1
        b = torch.rand(n, output_nc, size, size, generator=g) * 2 - 1
    else:
        a = torch.sign(torch.rand(n, 1, size, size, generator=g) - 0.9).expand(n, input_nc, size, size).contiguous()
        lo = torch.randn(n, output_nc, size // 8, size // 8, generator=g)
        b = F.interpolate(lo, size=(size, size), mode="bilinear", align_corners=False)
        b = b / b.abs().amax(dim=(1, 2, 3), keepdim=True).clamp_min(1e-6)
    z = torch.randn(n, nlatent, 1, 1, generator=g)
    return a.float(), b.float(), z.float()
