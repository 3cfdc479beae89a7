"""Evaluation inner loops of the reference, Python 3, running on the drop-in models:
``eval_mse_A`` (/root/reference/augmented_cyclegan/evaluate.py:10-19) and ``variational_ubo`` / ``eval_ubo_B``
(evaluate.py:21-148) -- per batch, `steps` iterations of: G_A_B forward on (real_A, z_B ~ q), Laplace log-likelihood of
real_B, KL to the prior, backward to (mu, logvar), RMSprop.  The network passes run through the fused plans
(model.predict_B stays differentiable with respect to z_B, networks._NetFn).  On the drop-in CUDA models the whole
iteration is fused (_variational_ubo_fused): dtg_ubo_laplace (objective + seed gradient), the generator backward to z
only (the reference's autograd also produces weight gradients nobody reads), dtg_ubo_latent_step (KLD, reparametrisation
backward, RMSprop, next z); the plain PyTorch objective below remains for any other model object (the reference's own, the
oracle adapter of the tests) and for compute_l1.
Visualisation (evaluate.py:79-86, 136-147) is left to the caller.  The reference hard-codes 64x64x3 in the
bits-per-pixel constant and the default logvar_B; here both follow real_B's shape (identical at 64x64x3).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def gauss_reparametrize(mu, logvar, n_sample=1):
    """model.py:15-22 (clamped to +-4)"""
    std = logvar.mul(0.5).exp()
    size = std.size()
    eps = std.detach().new_empty(size[0], n_sample, size[1]).normal_()
    z = eps.mul(std[:, None, :]).add(mu[:, None, :])
    z = torch.clamp(z, -4., 4.)
    return z.view(z.size(0) * z.size(1), z.size(2), 1, 1)


def log_prob_laplace(z, mu, log_var):
    """model.py:24-28"""
    sd = torch.exp(0.5 * log_var)
    return -0.5 * log_var - (torch.abs(z - mu) / sd) - float(np.log(2))


def kld_std_guss(mu, log_var):
    """model.py:45-53"""
    return -0.5 * torch.sum(log_var + 1. - mu ** 2 - torch.exp(log_var), dim=1)


def eval_mse_A(dataset, model, device="cuda"):
    """evaluate.py:10-19"""
    mse_A = []
    for batch in dataset:
        real_A, real_B = batch['A'].float().to(device), batch['B'].float().to(device)
        with torch.no_grad():
            pred_A = model.predict_A(real_B)
        mse_A.append(float(F.mse_loss(pred_A, real_A)))
    return np.mean(mse_A)


def _fused_model(model, real_A):
    """True for the drop-in models of dtg_b200.model on a CUDA device: variational_ubo then runs entirely on the C ABI"""
    try:
        from .model import _FusedCycleModel
    except Exception:       # pragma: no cover
        return False
    return (isinstance(model, _FusedCycleModel) and real_A.is_cuda and not model.opt.stoch_enc
            and not getattr(model, "ignore_noise", False))


def _variational_ubo_fused(model, real_A, real_B, steps, logvar_B, mu, logvar, verbose, q_out=None):
    """The loop of evaluate.py:89-124 on the fused kernels: per iteration ONE G_A_B forward, dtg_ubo_laplace (objective
    + seed gradient), G_A_B backward to z only (no weight gradients: the reference's autograd computes and drops
    them), dtg_ubo_latent_step (KLD, gradients through the reparametrisation, RMSprop, next z).  Nothing is read back
    before the loop ends (unless verbose).  The random draws are the reference's, in the reference's order."""
    from . import ops
    dev = real_A.device
    n, _, h, w = real_A.shape
    nz = model.opt.nlatent
    ex = model.netG_A_B._exec()
    c = ex.new_ctx(n, h, w, "ubo")
    i_out = model._head_idx(ex, "out")
    ops.pack_nchw(real_A.contiguous(), c.acts[0], 0)
    mu, logvar = mu.detach().clone().contiguous(), logvar.detach().clone().contiguous()
    sq_mu, sq_lv = torch.zeros_like(mu), torch.zeros_like(logvar)
    scal = torch.zeros(4, dtype=torch.float32, device=dev)
    ws = torch.zeros(1024, dtype=torch.float32, device=dev)
    lvb = logvar_B.to(dev).float().expand(1, *real_B.shape[1:]).contiguous()
    real_B = real_B.contiguous()
    ndim = real_B[0].numel()
    draw = lambda: mu.new_empty(n, 1, nz).normal_().view(n, nz)          # gauss_reparametrize's eps (model.py:19)
    eps = draw()
    z = torch.clamp(mu + eps * logvar.mul(0.5).exp(), -4., 4.)
    ubo_val = kld_val = bpp = None

    def read():
        logp, kld = scal[:2].tolist()
        u = -logp + kld + ndim * math.log(127.5)
        return u, kld, u / (ndim * math.log(2.))

    for i in range(steps):
        c.z.copy_(z)
        fake = ex.forward(c)["out"]
        ops.ubo_laplace(fake, real_B, lvb, scal, 0, c.dyraw[i_out], ws)
        ex.backward(c, {"out": True}, want_dx=False, want_dw=False, want_dz=True)
        eps_next = draw()
        ops.ubo_latent_step(mu, logvar, sq_mu, sq_lv, eps, eps_next, c.dz, 1e-2, 0.99, 1e-8, z, scal, 1)
        eps = eps_next
        if verbose:
            ubo_val, kld_val, bpp = read()
            print('[%d] UBO: %.4f, KLD: %.4f, BPP: %.4f' % (i, ubo_val, kld_val, bpp))
    if steps > 0 and not verbose:
        ubo_val, kld_val, bpp = read()
    if q_out is not None:
        q_out.update(mu=mu, logvar=logvar)
    return ubo_val, kld_val, bpp


def variational_ubo(model, real_A, real_B, steps, logvar_B=None, compute_l1=False, verbose=False, q_out=None):
    """evaluate.py:39-148 without the PNG dumps.  real_A / real_B live on the model's device.  Returns
    (ubo, kld, bpp) of the LAST evaluated iterate, like the reference.  On the drop-in models (CUDA) the whole loop
    runs on the fused kernels (_variational_ubo_fused); any other model object (e.g. the reference's, on CPU) takes
    the plain PyTorch restatement below."""
    dev = real_A.device
    nz = model.opt.nlatent
    dequant = torch.zeros(*real_B.size()).uniform_(0, 1. / 127.5).to(dev)                 # :43
    size = real_A.size()
    mu = torch.zeros(size[0], nz, device=dev, requires_grad=True)                         # :48-50
    logvar = torch.zeros(size[0], nz).fill_(math.log(0.01)).to(dev).requires_grad_(True)
    if logvar_B is None:
        logvar_B = torch.zeros(1, *real_B.shape[1:]).fill_(math.log(0.01)).to(dev)        # :52 ([1,3,64,64] there)
    if hasattr(model, 'netE_B'):                                                          # :56-62
        params = model.predict_enc_params(real_A, real_B)
        mu = params[0].detach().clone().requires_grad_(True)
        if len(params) == 2:
            logvar = params[1].detach().clone().requires_grad_(True)
    real_B = real_B + dequant                                                             # :67
    if _fused_model(model, real_A) and not compute_l1:
        return _variational_ubo_fused(model, real_A, real_B, steps, logvar_B, mu, logvar, verbose, q_out)
    iterative_opt = torch.optim.RMSprop([mu, logvar], lr=1e-2)                            # :65
    z_B = gauss_reparametrize(mu, logvar)                                                 # :70-71
    fake_B = model.predict_B(real_A, z_B)
    rec_B = None
    if compute_l1:                                                                        # :73-78
        with torch.no_grad():
            rec_B = fake_B.detach() if model.opt.stoch_enc else model.predict_B(real_A, mu.detach().view(size[0], nz, 1, 1))
    ubo_val = kld_val = bpp = None
    ndim = real_B[0].numel()          # the reference hard-codes 64 * 64 * 3 (:103, :106); generalised for N3's larger grids
    for i in range(steps):                                                                # :89
        log_prob = log_prob_laplace(real_B, fake_B, logvar_B).view(size[0], -1).sum(1)    # :93-94
        kld = kld_std_guss(mu, logvar)                                                    # :101
        ubo = (-log_prob + kld) + ndim * math.log(127.5)                                  # :103
        ubo_val = float(ubo.detach().mean(0))
        kld_val = float(kld.detach().mean(0))
        bpp = ubo_val / (ndim * math.log(2.))                                             # :106
        if verbose:
            msg = '[%d] UBO: %.4f, KLD: %.4f, BPP: %.4f' % (i, ubo_val, kld_val, bpp)
            if compute_l1:
                msg = '%s, L1: %.4f' % (msg, float(F.l1_loss(real_B, rec_B)))
            print(msg)
        loss = ubo.mean(0)                                                                # :118-121
        iterative_opt.zero_grad()
        loss.backward()
        iterative_opt.step()
        # :123-124; the iterate after the last RMSprop step is never scored nor back-propagated (the reference computes
        # and drops it), so it is not built as a differentiable forward
        with torch.set_grad_enabled(i + 1 < steps):
            z_B = gauss_reparametrize(mu, logvar)
            fake_B = model.predict_B(real_A, z_B)
        if compute_l1:
            with torch.no_grad():
                rec_B = fake_B.detach() if model.opt.stoch_enc else model.predict_B(real_A, mu.detach().view(size[0], nz, 1, 1))
    if q_out is not None:       # the optimised q(z) parameters (the reference drops them; tests compare them)
        q_out.update(mu=mu.detach(), logvar=logvar.detach())
    return ubo_val, kld_val, bpp


def eval_ubo_B(dataset, model, steps=500, logvar_B=None, compute_l1=False, verbose=False, device="cuda"):
    """evaluate.py:21-37"""
    ubo_B, bpp_B, kld_B = [], [], []
    for batch in dataset:
        real_A, real_B = batch['A'].float().to(device), batch['B'].float().to(device)
        ubo, kld, bpp = variational_ubo(model, real_A, real_B, steps, logvar_B, compute_l1, verbose)
        ubo_B.append(ubo)
        bpp_B.append(bpp)
        kld_B.append(kld)
    return np.mean(ubo_B), np.mean(bpp_B), np.mean(kld_B)
