"""Oracle restatement of the reference networks as pure functions over state dicts.

Test infrastructure only.  Every function takes ``p``: a dict keyed exactly like the
reference module's ``state_dict()`` (unique keys; the ``model.1x.{1,4,5}.*`` aliases
created at modules.py:145-146 are not needed) and returns what the reference
``forward`` returns.  BatchNorm layers run in training mode and update the running
statistics stored in ``p`` in place (the reference never calls ``.eval()``,
SURVEY.md section 9.4).  Paths cite /root/reference/augmented_cyclegan/.
"""
import torch
import torch.nn.functional as F

from .functional import cond_instance_norm, instance_norm


def _rpad(x, k):
    return F.pad(x, (k, k, k, k), mode="reflect")


def _cin(p, pre, x, z):
    return cond_instance_norm(x, z, p[pre + ".scale_conv.0.weight"], p[pre + ".scale_conv.0.bias"],
                              p[pre + ".shift_conv.0.weight"], p[pre + ".shift_conv.0.bias"])


def _in(p, pre, x):
    return instance_norm(x, p[pre + ".scale"], p[pre + ".shift"])


def _n_blocks(p):
    """res-blocks present in a generator state dict: 3 for every reference model (networks.py:173, 225 ignore n_blocks);
    other counts are the N3 extension with model indices 10 .. 9+n and the tail shifted accordingly"""
    n = 0
    while "model.%d.conv_block.4.weight" % (10 + n) in p:
        n += 1
    return n


def cin_resnet_generator(p, x, z, capture=None):
    """networks.py:149-197 (3 CIN res-blocks regardless of n_blocks, networks.py:173)."""
    nb = _n_blocks(p)
    t0 = 10 + nb
    def cap(name, t):
        if capture is not None:
            capture[name] = t
        return t
    h = F.conv2d(_rpad(x, 3), p["model.1.weight"], p["model.1.bias"])
    h = cap("model.3", F.relu(_cin(p, "model.2", cap("model.1", h), z)))
    h = F.conv2d(h, p["model.4.weight"], p["model.4.bias"], padding=1)
    h = cap("model.6", F.relu(_cin(p, "model.5", cap("model.4", h), z)))
    h = F.conv2d(h, p["model.7.weight"], p["model.7.bias"], stride=2, padding=1)
    h = cap("model.9", F.relu(_cin(p, "model.8", cap("model.7", h), z)))
    for i in range(10, t0):  # modules.py:139-188
        b = "model.%d.conv_block" % i
        t = F.conv2d(_rpad(h, 1), p[b + ".1.module1.weight"], p[b + ".1.module1.bias"])
        t = F.relu(_cin(p, b + ".1.module2", t, z))
        t = F.conv2d(_rpad(t, 1), p[b + ".4.weight"], p[b + ".4.bias"])
        t = _in(p, b + ".5", t)
        h = cap("model.%d" % i, F.relu(h + t))
    k = lambda j: "model.%d" % (t0 + j)
    h = F.conv_transpose2d(h, p[k(0) + ".weight"], p[k(0) + ".bias"], stride=2, padding=1, output_padding=1)
    h = cap(k(2), F.relu(_cin(p, k(1), cap(k(0), h), z)))
    h = F.conv2d(h, p[k(3) + ".weight"], p[k(3) + ".bias"], padding=1)
    h = cap(k(5), F.relu(_cin(p, k(4), cap(k(3), h), z)))
    h = F.conv2d(h, p[k(6) + ".weight"], p[k(6) + ".bias"], padding=3)
    return torch.tanh(cap(k(6), h))


def resnet_generator(p, x, capture=None):
    """networks.py:203-252; res-block = conv-ReLU-conv-IN, relu(x+.) (modules.py:193-235)."""
    nb = _n_blocks(p)
    t0 = 10 + nb
    def cap(name, t):
        if capture is not None:
            capture[name] = t
        return t
    h = F.conv2d(_rpad(x, 3), p["model.1.weight"], p["model.1.bias"])
    h = cap("model.3", F.relu(_in(p, "model.2", cap("model.1", h))))
    h = F.conv2d(h, p["model.4.weight"], p["model.4.bias"], padding=1)
    h = cap("model.6", F.relu(_in(p, "model.5", cap("model.4", h))))
    h = F.conv2d(h, p["model.7.weight"], p["model.7.bias"], stride=2, padding=1)
    h = cap("model.9", F.relu(_in(p, "model.8", cap("model.7", h))))
    for i in range(10, t0):
        b = "model.%d.conv_block" % i
        t = F.relu(F.conv2d(_rpad(h, 1), p[b + ".1.weight"], p[b + ".1.bias"]))
        t = F.conv2d(_rpad(t, 1), p[b + ".4.weight"], p[b + ".4.bias"])
        t = _in(p, b + ".5", t)
        h = cap("model.%d" % i, F.relu(h + t))
    k = lambda j: "model.%d" % (t0 + j)
    h = F.conv_transpose2d(h, p[k(0) + ".weight"], p[k(0) + ".bias"], stride=2, padding=1, output_padding=1)
    h = cap(k(2), F.relu(_in(p, k(1), cap(k(0), h))))
    h = F.conv2d(h, p[k(3) + ".weight"], p[k(3) + ".bias"], padding=1)
    h = cap(k(5), F.relu(_in(p, k(4), cap(k(3), h))))
    h = F.conv2d(h, p[k(6) + ".weight"], p[k(6) + ".bias"], padding=3)
    return torch.tanh(cap(k(6), h))


def cin_resnet_block(p, x, z):
    """modules.py:139-188 standalone: relu(x + IN(conv(rpad(relu(CIN(conv(rpad(x)), z)))))); keys of the block's own
    state_dict (conv_block.1.module1.*, conv_block.1.module2.*, conv_block.4.*, conv_block.5.*)"""
    b = "conv_block"
    t = F.conv2d(_rpad(x, 1), p[b + ".1.module1.weight"], p[b + ".1.module1.bias"])
    t = F.relu(_cin(p, b + ".1.module2", t, z))
    t = F.conv2d(_rpad(t, 1), p[b + ".4.weight"], p[b + ".4.bias"])
    return F.relu(x + _in(p, b + ".5", t))


def resnet_block(p, x):
    """modules.py:193-235 standalone: relu(x + IN(conv(rpad(relu(conv(rpad(x)))))))"""
    b = "conv_block"
    t = F.relu(F.conv2d(_rpad(x, 1), p[b + ".1.weight"], p[b + ".1.bias"]))
    t = F.conv2d(_rpad(t, 1), p[b + ".4.weight"], p[b + ".4.bias"])
    return F.relu(x + _in(p, b + ".5", t))


def discriminator(p, x, capture=None):
    """networks.py:308-349 PatchGAN, 4x4 kernels, strides 2,2,1,1,1, pad 1."""
    h = F.leaky_relu(F.conv2d(x, p["model.0.weight"], p["model.0.bias"], stride=2, padding=1), 0.2)
    for ci, ni, s in ((2, 3, 2), (5, 6, 1), (8, 9, 1)):
        h = F.conv2d(h, p["model.%d.weight" % ci], p["model.%d.bias" % ci], stride=s, padding=1)
        h = F.leaky_relu(_in(p, "model.%d" % ni, h), 0.2)
        if capture is not None:
            capture["model.%d" % (ni + 1)] = h
    return F.conv2d(h, p["model.11.weight"], p["model.11.bias"], stride=1, padding=1)


def discriminator_edges(p, x, capture=None):
    """networks.py:352-393: four 3x3 stride-2 convs then a 4x4 valid conv."""
    h = F.leaky_relu(F.conv2d(x, p["model.0.weight"], p["model.0.bias"], stride=2, padding=1), 0.2)
    for ci, ni in ((2, 3), (5, 6), (8, 9)):
        h = F.conv2d(h, p["model.%d.weight" % ci], p["model.%d.bias" % ci], stride=2, padding=1)
        h = F.leaky_relu(_in(p, "model.%d" % ni, h), 0.2)
        if capture is not None:
            capture["model.%d" % (ni + 1)] = h
    return F.conv2d(h, p["model.11.weight"], p["model.11.bias"])


def _bn(p, pre, x, momentum=0.1, eps=1e-5):
    """nn.BatchNorm{1,2}d in training mode (networks.py:25,407-415,450-462)."""
    y = F.batch_norm(x, p[pre + ".running_mean"], p[pre + ".running_var"], p[pre + ".weight"],
                     p[pre + ".bias"], True, momentum, eps)
    p[pre + ".num_batches_tracked"] += 1
    return y


def discriminator_latent(p, z):
    """networks.py:396-433."""
    h = z.reshape(z.shape[0], -1)
    for li, bi in ((0, 1), (3, 4), (6, 7)):
        h = F.linear(h, p["model.%d.weight" % li], p["model.%d.bias" % li])
        h = F.leaky_relu(_bn(p, "model.%d" % bi, h), 0.2)
    return F.linear(h, p["model.9.weight"], p["model.9.bias"])


def latent_encoder(p, x):
    """networks.py:438-482; returns (mu, logvar) flattened to [N, -1].

    N3 extension: a state dict with extra stride-2 stages (conv 8nef->8nef k3 s2 p1 no bias + BN + ReLU at
    conv_modules.11, 14, ... before the 4x4 valid conv) is the encoder for 64 * 2^k inputs; with none it is the reference's."""
    h = F.relu(F.conv2d(x, p["conv_modules.0.weight"], p["conv_modules.0.bias"], stride=2, padding=1))
    ci = 2
    while p["conv_modules.%d.weight" % ci].shape[-1] == 3:
        h = F.conv2d(h, p["conv_modules.%d.weight" % ci], None, stride=2, padding=1)
        h = F.relu(_bn(p, "conv_modules.%d" % (ci + 1), h))
        ci += 3
    h = F.conv2d(h, p["conv_modules.%d.weight" % ci], None)
    h = F.relu(_bn(p, "conv_modules.%d" % (ci + 1), h))
    mu = F.conv2d(h, p["enc_mu.weight"], p["enc_mu.bias"])
    lv = F.conv2d(h, p["enc_logvar.weight"], p["enc_logvar.bias"])
    return mu.reshape(mu.shape[0], -1), lv.reshape(lv.shape[0], -1)


# --------------------------------------------------------------------------------------
# Parameter construction with the reference's init *distributions* (networks.py:13-21,
# modules.py:78-81; nn.Linear / BatchNorm1d keep torch defaults).  RNG streams differ from
# the reference constructors, so parity tests always copy one state dict into both sides.
# --------------------------------------------------------------------------------------

def _conv(p, name, cout, cin, k, g, bias=True):
    p[name + ".weight"] = torch.empty(cout, cin, k, k).normal_(0.0, 0.02, generator=g)
    if bias:
        p[name + ".bias"] = torch.zeros(cout)


def _inorm(p, name, c, g):
    p[name + ".scale"] = torch.empty(c).normal_(0.0, 0.02, generator=g)
    p[name + ".shift"] = torch.zeros(c)


def _cinorm(p, name, c, nz, g):
    for br in ("shift_conv", "scale_conv"):
        _conv(p, "%s.%s.0" % (name, br), c, nz, 1, g)


def _bnorm(p, name, c, g, two_d):
    p[name + ".weight"] = torch.empty(c).normal_(1.0, 0.02, generator=g) if two_d else torch.ones(c)
    p[name + ".bias"] = torch.zeros(c)
    p[name + ".running_mean"] = torch.zeros(c)
    p[name + ".running_var"] = torch.ones(c)
    p[name + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)


def init_generator(g, in_nc, out_nc, ngf, nlatent=None, n_blocks=3):
    """nlatent=None -> ResnetGenerator keys, else CINResnetGenerator keys.  n_blocks != 3 is the N3 extension."""
    p = {}
    cin = nlatent is not None
    def norm(name, c):
        _cinorm(p, name, c, nlatent, g) if cin else _inorm(p, name, c, g)
    _conv(p, "model.1", ngf, in_nc, 7, g); norm("model.2", ngf)
    _conv(p, "model.4", 2 * ngf, ngf, 3, g); norm("model.5", 2 * ngf)
    _conv(p, "model.7", 4 * ngf, 2 * ngf, 3, g); norm("model.8", 4 * ngf)
    t0 = 10 + n_blocks
    for i in range(10, t0):
        b = "model.%d.conv_block" % i
        if cin:
            _conv(p, b + ".1.module1", 4 * ngf, 4 * ngf, 3, g)
            _cinorm(p, b + ".1.module2", 4 * ngf, nlatent, g)
        else:
            _conv(p, b + ".1", 4 * ngf, 4 * ngf, 3, g)
        _conv(p, b + ".4", 4 * ngf, 4 * ngf, 3, g)
        _inorm(p, b + ".5", 4 * ngf, g)
    # ConvTranspose2d weight layout [Cin, Cout, k, k] (networks.py:178-179)
    p["model.%d.weight" % t0] = torch.empty(4 * ngf, 2 * ngf, 3, 3).normal_(0.0, 0.02, generator=g)
    p["model.%d.bias" % t0] = torch.zeros(2 * ngf)
    norm("model.%d" % (t0 + 1), 2 * ngf)
    _conv(p, "model.%d" % (t0 + 3), ngf, 2 * ngf, 3, g); norm("model.%d" % (t0 + 4), ngf)
    _conv(p, "model.%d" % (t0 + 6), out_nc, ngf, 7, g)
    return p


def init_discriminator(g, in_nc, ndf, edges):
    p = {}
    k = 3 if edges else 4
    _conv(p, "model.0", ndf, in_nc, k, g)
    _conv(p, "model.2", 2 * ndf, ndf, k, g); _inorm(p, "model.3", 2 * ndf, g)
    _conv(p, "model.5", 4 * ndf, 2 * ndf, k, g); _inorm(p, "model.6", 4 * ndf, g)
    _conv(p, "model.8", 4 * ndf, 4 * ndf, k, g); _inorm(p, "model.9", 4 * ndf, g)
    _conv(p, "model.11", 1, 4 * ndf, 4, g)
    return p


def init_discriminator_latent(g, nlatent, ndf):
    p = {}
    dims = [(0, nlatent, ndf), (3, ndf, ndf), (6, ndf, ndf), (9, ndf, 1)]
    for li, fi, fo in dims:
        bound = 1.0 / fi ** 0.5  # nn.Linear default (kaiming_uniform a=sqrt(5)) -> U(+-1/sqrt(fan_in))
        p["model.%d.weight" % li] = (torch.rand(fo, fi, generator=g) * 2 - 1) * bound
        p["model.%d.bias" % li] = (torch.rand(fo, generator=g) * 2 - 1) * bound
    for bi in (1, 4, 7):
        _bnorm(p, "model.%d" % bi, ndf, g, two_d=False)
    return p


def init_encoder(g, nlatent, in_nc, nef, img_size=64):
    """img_size = 64 * 2^k adds k stride-2 stages before the 4x4 valid conv (N3 extension; 64 = the reference)"""
    p = {}
    _conv(p, "conv_modules.0", nef, in_nc, 3, g)
    _conv(p, "conv_modules.2", 2 * nef, nef, 3, g, bias=False); _bnorm(p, "conv_modules.3", 2 * nef, g, True)
    _conv(p, "conv_modules.5", 4 * nef, 2 * nef, 3, g, bias=False); _bnorm(p, "conv_modules.6", 4 * nef, g, True)
    _conv(p, "conv_modules.8", 8 * nef, 4 * nef, 3, g, bias=False); _bnorm(p, "conv_modules.9", 8 * nef, g, True)
    ci = 11
    while img_size > 64:
        _conv(p, "conv_modules.%d" % ci, 8 * nef, 8 * nef, 3, g, bias=False); _bnorm(p, "conv_modules.%d" % (ci + 1), 8 * nef, g, True)
        ci += 3
        img_size //= 2
    _conv(p, "conv_modules.%d" % ci, 8 * nef, 8 * nef, 4, g, bias=False); _bnorm(p, "conv_modules.%d" % (ci + 1), 8 * nef, g, True)
    _conv(p, "enc_mu", nlatent, 8 * nef, 1, g)
    _conv(p, "enc_logvar", nlatent, 8 * nef, 1, g)
    return p


NET_NAMES = ("netG_A_B", "netG_B_A", "netE_B", "netD_A", "netD_B", "netD_z_B")


def init_model_state(seed=1234, input_nc=3, output_nc=3, ngf=32, nef=32, ndf=64, nlatent=16,
                     enc_A_B=True, perturb=0.0, n_blocks=3, img_size=64):
    """Parameters of the six networks AugmentedCycleGAN.__init__ builds (model.py:348-376).

    perturb > 0 adds N(0, perturb) to every bias / shift so tests exercise non-zero values.
    """
    g = torch.Generator().manual_seed(seed)
    st = {
        "netG_A_B": init_generator(g, input_nc, output_nc, ngf, nlatent, n_blocks),
        "netG_B_A": init_generator(g, output_nc, input_nc, ngf, None, n_blocks),
        "netE_B": init_encoder(g, nlatent, output_nc + (input_nc if enc_A_B else 0), nef, img_size),
        "netD_A": init_discriminator(g, input_nc, 32, edges=True),   # ndf hard-coded, model.py:367
        "netD_B": init_discriminator(g, output_nc, ndf, edges=False),
        "netD_z_B": init_discriminator_latent(g, nlatent, ndf),
    }
    if perturb > 0:
        for sd in st.values():
            for k, v in sd.items():
                if k.endswith(".bias") or k.endswith(".shift"):
                    v.add_(torch.empty_like(v).normal_(0.0, perturb, generator=g))
    return st


def is_noise_grad(net_name, key, n_blocks=3):
    """True for biases whose exact gradient is 0 because a mean-removing norm follows the layer
    (instance / conditional-instance / batch norm): the reference's autograd produces pure fp32
    rounding noise there (|g| ~ 1e-6 of the weight gradient), so parity is checked against a noise
    floor instead of relatively."""
    if not key.endswith(".bias"):
        return False
    k = key[:-5]
    if net_name in ("netG_A_B", "netG_B_A"):
        if k in ("model.1", "model.4", "model.7", "model.%d" % (10 + n_blocks), "model.%d" % (13 + n_blocks)):
            return True
        if k.endswith("conv_block.4") or k.endswith("conv_block.1.module1"):
            return True
        return False
    if net_name in ("netD_A", "netD_B"):
        return k in ("model.2", "model.5", "model.8")
    if net_name == "netD_z_B":
        return k in ("model.0", "model.3", "model.6")
    return False
