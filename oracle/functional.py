"""Oracle restatement of the reference's norm / loss primitives (test infrastructure).

Each function cites the reference lines it follows (paths relative to
/root/reference/augmented_cyclegan/).
"""
import math

import torch
import torch.nn.functional as F


def instance_norm(x, scale=None, shift=None, eps=1e-5):
    """modules.py:83-97 -- per-(n,c) mean, BIASED variance (mean of centred squares)."""
    n, c, h, w = x.shape
    flat = x.reshape(n, c, h * w)
    mu = flat.mean(2, keepdim=True)
    xc = flat - mu
    r = torch.rsqrt((xc * xc).mean(2, keepdim=True) + eps)
    y = (xc * r).reshape(n, c, h, w)
    if scale is not None:
        y = y * scale[:, None, None] + shift[:, None, None]
    return y


def cin_affine(z, w_scale, b_scale, w_shift, b_shift):
    """modules.py:111-118,123-124 -- relu(1x1 conv(z)) for scale and shift; z is [N,Z,1,1]."""
    scale = F.relu(F.conv2d(z, w_scale, b_scale))
    shift = F.relu(F.conv2d(z, w_shift, b_shift))
    return scale, shift


def cond_instance_norm(x, z, w_scale, b_scale, w_shift, b_shift, eps=1e-5):
    """modules.py:120-132 -- per-(n,c) mean, UNBIASED variance (torch.var default)."""
    scale, shift = cin_affine(z, w_scale, b_scale, w_shift, b_shift)
    n, c, h, w = x.shape
    flat = x.reshape(n, c, h * w)
    mu = flat.mean(2, keepdim=True)
    var = flat.var(2, keepdim=True)
    y = ((flat - mu) * torch.rsqrt(var + eps)).reshape(n, c, h, w)
    return y * scale + shift


def lsgan(pred, target_is_real):
    """model.py:56-72 with use_sigmoid=False: MSE against a constant 1 / 0 map."""
    t = torch.ones_like(pred) if target_is_real else torch.zeros_like(pred)
    return F.mse_loss(pred, t)


def kld_std_gauss(mu, log_var):
    """model.py:45-53."""
    return -0.5 * torch.sum(log_var + 1.0 - mu ** 2 - torch.exp(log_var), dim=1)


def log_prob_gaussian(z, mu, log_var):
    """model.py:31-34."""
    res = -0.5 * log_var - ((z - mu) ** 2.0 / (2.0 * torch.exp(log_var)))
    return res - 0.5 * math.log(2 * math.pi)


def gauss_reparametrize(mu, logvar, n_sample=1):
    """model.py:15-22 (consumes RNG; only on the stoch_enc branch)."""
    std = (0.5 * logvar).exp()
    eps = torch.randn(std.shape[0], n_sample, std.shape[1], dtype=std.dtype, device=std.device)
    z = (eps * std[:, None, :] + mu[:, None, :]).clamp(-4.0, 4.0)
    return z.reshape(z.shape[0] * z.shape[1], z.shape[2], 1, 1)
