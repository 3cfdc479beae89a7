"""Networks with the reference's class names, constructor signatures, attribute names and state_dict
keys (/root/reference/augmented_cyclegan/networks.py:13-482), executed by fused sm_100a plans.

Every class builds the SAME nn.Module tree as the reference (so ``state_dict()`` keys, including the
``model.1x.{1,4,5}.*`` aliases, are identical and checkpoints interchange), then describes its forward
as an engine plan.  ``forward`` takes / returns fp32 NCHW tensors like the reference and is
differentiable through torch autograd (parameter gradients accumulate into ``.grad``); the fused
training step (model.py) drives the same plans directly.

``gpu_ids`` is accepted for signature compatibility; multi-GPU is one process per GPU (parallel.py),
not ``nn.parallel.data_parallel``.
"""
import functools
import weakref

import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .engine import Layer, NetExec, ParamArena
from .modules import (CINResnetBlock, CondInstanceNorm, InstanceNorm, InstanceNorm2d, ResnetBlock,  # noqa: F401
                      TwoInputSequential)


def weights_init(m):
    """networks.py:13-21"""
    classname = m.__class__.__name__
    if classname.find('Conv') != -1:
        m.weight.data.normal_(0.0, 0.02)
        if hasattr(m.bias, 'data'):
            m.bias.data.fill_(0)
    elif classname.find('BatchNorm2d') != -1:
        m.weight.data.normal_(1.0, 0.02)
        m.bias.data.fill_(0)


def get_norm_layer(norm_type='instance'):
    """networks.py:23-30"""
    if norm_type == 'batch':
        return functools.partial(nn.BatchNorm2d, affine=True)
    if norm_type == 'instance':
        return functools.partial(InstanceNorm2d, affine=True)
    raise NotImplementedError('normalization layer [%s] is not found' % norm_type)


def _norm_kind(m):
    if isinstance(m, CondInstanceNorm):
        return L.NORM_COND_INSTANCE
    if isinstance(m, InstanceNorm):
        return L.NORM_INSTANCE
    if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)):
        return L.NORM_BATCH
    raise NotImplementedError("dtg_b200: unsupported norm layer %s" % type(m).__name__)


def _release_ctx(c):
    c.busy = False


class _NetFn(torch.autograd.Function):
    """Differentiable network call: forward / backward run the fused plan on a private context."""

    @staticmethod
    def forward(ctx, net, head_names, x, z, *extra):
        ex = net._exec()
        n, _, h, w = x.shape
        ctx.set_materialize_grads(False)
        ex.repack()                       # weights may have been changed by any torch optimizer
        slot = 0
        while getattr(ex.new_ctx(n, h, w, tag=("autograd", slot)), "busy", False):
            slot += 1
        c = ex.new_ctx(n, h, w, tag=("autograd", slot))
        # the activation context stays reserved until backward() has consumed it -- or until the autograd node dies
        # without ever being back-propagated (a grad-enabled forward whose output is dropped), else it would leak
        c.busy = torch.is_grad_enabled()
        if c.busy:
            weakref.finalize(ctx, _release_ctx, c)
        srcs = (x,) + tuple(extra)
        off = 0
        for s in srcs:
            ops.pack_nchw(s.detach().contiguous().float(), c.acts[0], off)
            off += s.shape[1]
        if z is not None:
            c.z.copy_(z.detach().reshape(n, -1))
        ex.forward(c)
        ctx.net, ctx.c, ctx.heads = net, c, head_names
        ctx.splits = [s.shape[1] for s in srcs]
        ctx.zshape = z.shape if z is not None else None
        outs = tuple(c.heads[hn].clone() for hn in head_names)
        ctx.save_for_backward(*outs)
        return outs if len(outs) > 1 else outs[0]

    @staticmethod
    def backward(ctx, *douts):
        net, c = ctx.net, ctx.c
        ex = net._exec()
        c.busy = False
        seeds = {}
        for hn, dy, y in zip(ctx.heads, douts, ctx.saved_tensors):
            if dy is None:
                continue
            idx = [i for i, ly in enumerate(ex.layers) if ly.name == hn][0]
            ly = ex.layers[idx]
            ops.pack_nchw(dy.contiguous().float(), c.dyraw[idx], 0, tanh_y=y if ly.act == L.ACT_TANH else None, reflect=False)
            seeds[hn] = True
        gin = ex.backward(c, seeds, want_dx=True, want_dw=True, want_dz=ctx.zshape is not None)
        grads, off = [], 0
        for cs in ctx.splits:
            d = torch.empty(c.n, cs, gin.h, gin.w, device=gin.t.device)
            ops.grad_gather([gin], [off], cs, out=None, out_nchw=d)
            grads.append(d)
            off += cs
        dz = c.dz.reshape(ctx.zshape).clone() if ctx.zshape is not None else None
        return (None, None, grads[0], dz) + tuple(grads[1:])


class _FusedNet(nn.Module):
    """Common plumbing: lazy arena / plan construction (after .cuda()), autograd entry."""

    _inactive = ()

    def _exec(self):
        ex = getattr(self, "_ex", None)
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("dtg_b200 networks run only on CUDA (sm_100a); there is no CPU fallback")
        if ex is None or ex.arena.device != dev:
            arena = ParamArena(self, inactive=self._inactive)
            ex = self._build_exec(arena)
            object.__setattr__(self, "_ex", ex)
        ex.prepare()
        return ex

    def _call(self, heads, x, z=None, *extra):
        return _NetFn.apply(self, heads, x, z, *extra)


def _res_block_count(n_blocks, honor_n_blocks):
    """The reference builds 3 res-blocks whatever n_blocks says (`for i in range(3)`, networks.py:173, 225).
    honor_n_blocks=True is the N3 extension: range(n_blocks)."""
    return int(n_blocks) if honor_n_blocks else 3


def _gen_layers(model, cin_mode, input_nc, output_nc, ngf, nb=3):
    """Layer plan shared by CINResnetGenerator / ResnetGenerator (networks.py:158-189, 210-244); nb res-blocks at
    model.10 .. model.(9+nb), the decoder tail follows at model.(10+nb)."""
    m = model
    relu = L.ACT_RELU
    nk = lambda mod: _norm_kind(mod)
    ls = [
        Layer("model.1", 0, m[1], input_nc, ngf, 7, 1, 3, norm=nk(m[2]), act=relu, norm_mod=m[2]),
        Layer("model.4", 1, m[4], ngf, 2 * ngf, 3, 1, 1, norm=nk(m[5]), act=relu, norm_mod=m[5]),
        Layer("model.7", 2, m[7], 2 * ngf, 4 * ngf, 3, 2, 1, norm=nk(m[8]), act=relu, norm_mod=m[8],
              out_halo=1 if nb > 0 else 0),
    ]
    src = 3
    for bi, idx in enumerate(range(10, 10 + nb)):
        cb = m[idx].conv_block
        last = bi == nb - 1
        if cin_mode:
            conv1, n1 = cb[1].module1, cb[1].module2
            ls.append(Layer("model.%d.a" % idx, src, conv1, 4 * ngf, 4 * ngf, 3, 1, 1, norm=nk(n1), act=relu,
                            norm_mod=n1, out_halo=1))
        else:
            ls.append(Layer("model.%d.a" % idx, src, cb[1], 4 * ngf, 4 * ngf, 3, 1, 1, norm=L.NORM_NONE, act=relu,
                            out_halo=1))
        ls.append(Layer("model.%d.b" % idx, src + 1, cb[4], 4 * ngf, 4 * ngf, 3, 1, 1, norm=nk(cb[5]), act=relu,
                        norm_mod=cb[5], out_halo=0 if last else 1, residual=src))
        src += 2
    t0 = 10 + nb
    ls += [
        Layer("model.%d" % t0, src, m[t0], 4 * ngf, 2 * ngf, 3, 2, 1, transposed=True, norm=nk(m[t0 + 1]), act=relu,
              norm_mod=m[t0 + 1]),
        Layer("model.%d" % (t0 + 3), src + 1, m[t0 + 3], 2 * ngf, ngf, 3, 1, 1, norm=nk(m[t0 + 4]), act=relu,
              norm_mod=m[t0 + 4]),
        Layer("out", src + 2, m[t0 + 6], ngf, output_nc, 7, 1, 3, act=L.ACT_TANH, head=True),
    ]
    return ls


class CINResnetGenerator(_FusedNet):
    """networks.py:149-197.  n_blocks is accepted and ignored exactly like the reference (3 blocks) unless
    honor_n_blocks=True (extension N3)."""

    def __init__(self, nlatent, input_nc, output_nc, ngf=64, norm_layer=CondInstanceNorm,
                 use_dropout=False, n_blocks=9, gpu_ids=[], padding_type='reflect', honor_n_blocks=False):
        assert (n_blocks >= 0)
        super().__init__()
        self.n_res = _res_block_count(n_blocks, honor_n_blocks)
        if use_dropout or padding_type != 'reflect' or norm_layer is not CondInstanceNorm:
            raise NotImplementedError("dtg_b200: only the reference's default generator configuration "
                                      "(CondInstanceNorm, reflect padding, no dropout) is implemented")
        self.gpu_ids = gpu_ids
        self.nlatent, self.input_nc, self.output_nc, self.ngf = nlatent, input_nc, output_nc, ngf
        model = [nn.ReflectionPad2d(3),
                 nn.Conv2d(input_nc, ngf, kernel_size=7, padding=0, stride=1, bias=True),
                 norm_layer(ngf, nlatent), nn.ReLU(True),
                 nn.Conv2d(ngf, 2 * ngf, kernel_size=3, padding=1, stride=1, bias=True),
                 norm_layer(2 * ngf, nlatent), nn.ReLU(True),
                 nn.Conv2d(2 * ngf, 4 * ngf, kernel_size=3, padding=1, stride=2, bias=True),
                 norm_layer(4 * ngf, nlatent), nn.ReLU(True)]
        for i in range(self.n_res):
            model += [CINResnetBlock(x_dim=4 * ngf, z_dim=nlatent, padding_type=padding_type,
                                     norm_layer=norm_layer, use_dropout=use_dropout, use_bias=True)]
        model += [nn.ConvTranspose2d(4 * ngf, 2 * ngf, kernel_size=3, stride=2, padding=1, output_padding=1, bias=True),
                  norm_layer(2 * ngf, nlatent), nn.ReLU(True),
                  nn.Conv2d(2 * ngf, ngf, kernel_size=3, padding=1, stride=1, bias=True),
                  norm_layer(ngf, nlatent), nn.ReLU(True),
                  nn.Conv2d(ngf, output_nc, kernel_size=7, padding=3), nn.Tanh()]
        self.model = TwoInputSequential(*model)

    def _build_exec(self, arena):
        return NetExec(self, _gen_layers(self.model, True, self.input_nc, self.output_nc, self.ngf, self.n_res),
                       self.input_nc, 3, arena, nz=self.nlatent)

    def forward(self, input, noise):
        return self._call(("out",), input, noise)


class ResnetGenerator(_FusedNet):
    """networks.py:203-252 (3 res-blocks like the reference unless honor_n_blocks=True, extension N3)."""

    def __init__(self, input_nc, output_nc, ngf=64, norm_layer=InstanceNorm2d, use_dropout=False,
                 n_blocks=9, gpu_ids=[], padding_type='reflect', honor_n_blocks=False):
        assert (n_blocks >= 0)
        super().__init__()
        self.n_res = _res_block_count(n_blocks, honor_n_blocks)
        if use_dropout or padding_type != 'reflect':
            raise NotImplementedError("dtg_b200: only reflect padding without dropout is implemented")
        self.gpu_ids = gpu_ids
        self.input_nc, self.output_nc, self.ngf = input_nc, output_nc, ngf
        model = [nn.ReflectionPad2d(3),
                 nn.Conv2d(input_nc, ngf, kernel_size=7, padding=0, stride=1, bias=True), norm_layer(ngf), nn.ReLU(True),
                 nn.Conv2d(ngf, 2 * ngf, kernel_size=3, padding=1, stride=1, bias=True), norm_layer(2 * ngf), nn.ReLU(True),
                 nn.Conv2d(2 * ngf, 4 * ngf, kernel_size=3, padding=1, stride=2, bias=True), norm_layer(4 * ngf),
                 nn.ReLU(True)]
        for i in range(self.n_res):
            model += [ResnetBlock(4 * ngf, padding_type=padding_type, norm_layer=norm_layer,
                                  use_dropout=use_dropout, use_bias=True)]
        model += [nn.ConvTranspose2d(4 * ngf, 2 * ngf, kernel_size=3, stride=2, padding=1, output_padding=1, bias=True),
                  norm_layer(2 * ngf), nn.ReLU(True),
                  nn.Conv2d(2 * ngf, ngf, kernel_size=3, padding=1, bias=True), norm_layer(ngf), nn.ReLU(True),
                  nn.Conv2d(ngf, output_nc, kernel_size=7, padding=3), nn.Tanh()]
        self.model = nn.Sequential(*model)

    def _build_exec(self, arena):
        return NetExec(self, _gen_layers(self.model, False, self.input_nc, self.output_nc, self.ngf, self.n_res),
                       self.input_nc, 3, arena)

    def forward(self, input):
        return self._call(("out",), input)


class Discriminator(_FusedNet):
    """networks.py:308-349 PatchGAN (D_B)."""

    def __init__(self, input_nc, ndf=64, norm_layer=nn.BatchNorm2d, use_sigmoid=False, gpu_ids=[]):
        super().__init__()
        if use_sigmoid:
            raise NotImplementedError("dtg_b200: LSGAN only (the reference's BCE branch is broken, SURVEY 3.4)")
        self.gpu_ids, self.input_nc, self.ndf = gpu_ids, input_nc, ndf
        kw = 4
        self.model = nn.Sequential(
            nn.Conv2d(input_nc, ndf, kernel_size=kw, stride=2, padding=1, bias=True), nn.LeakyReLU(0.2, True),
            nn.Conv2d(ndf, 2 * ndf, kernel_size=kw, stride=2, padding=1, bias=True), norm_layer(2 * ndf), nn.LeakyReLU(0.2, True),
            nn.Conv2d(2 * ndf, 4 * ndf, kernel_size=kw, stride=1, padding=1, bias=True), norm_layer(4 * ndf), nn.LeakyReLU(0.2, True),
            nn.Conv2d(4 * ndf, 4 * ndf, kernel_size=kw, stride=1, padding=1, bias=True), norm_layer(4 * ndf), nn.LeakyReLU(0.2, True),
            nn.Conv2d(4 * ndf, 1, kernel_size=kw, stride=1, padding=1))

    def _build_exec(self, arena):
        m, ndf, lr = self.model, self.ndf, L.ACT_LRELU
        ls = [Layer("model.0", 0, m[0], self.input_nc, ndf, 4, 2, 1, act=lr),
              Layer("model.2", 1, m[2], ndf, 2 * ndf, 4, 2, 1, norm=_norm_kind(m[3]), act=lr, norm_mod=m[3]),
              Layer("model.5", 2, m[5], 2 * ndf, 4 * ndf, 4, 1, 1, norm=_norm_kind(m[6]), act=lr, norm_mod=m[6]),
              Layer("model.8", 3, m[8], 4 * ndf, 4 * ndf, 4, 1, 1, norm=_norm_kind(m[9]), act=lr, norm_mod=m[9]),
              Layer("out", 4, m[11], 4 * ndf, 1, 4, 1, 1, head=True)]
        return NetExec(self, ls, self.input_nc, 0, arena)

    def forward(self, input):
        return self._call(("out",), input)


NLayerDiscriminator = Discriminator   # north-star alias; the reference class is `Discriminator`


class Discriminator_edges(_FusedNet):
    """networks.py:352-393 (D_A)."""

    def __init__(self, input_nc, ndf=64, norm_layer=nn.BatchNorm2d, use_sigmoid=False, gpu_ids=[]):
        super().__init__()
        if use_sigmoid:
            raise NotImplementedError("dtg_b200: LSGAN only")
        self.gpu_ids, self.input_nc, self.ndf = gpu_ids, input_nc, ndf
        kw = 3
        self.model = nn.Sequential(
            nn.Conv2d(input_nc, ndf, kernel_size=kw, stride=2, padding=1, bias=True), nn.LeakyReLU(0.2, True),
            nn.Conv2d(ndf, 2 * ndf, kernel_size=kw, stride=2, padding=1, bias=True), norm_layer(2 * ndf), nn.LeakyReLU(0.2, True),
            nn.Conv2d(2 * ndf, 4 * ndf, kernel_size=kw, stride=2, padding=1, bias=True), norm_layer(4 * ndf), nn.LeakyReLU(0.2, True),
            nn.Conv2d(4 * ndf, 4 * ndf, kernel_size=kw, stride=2, padding=1, bias=True), norm_layer(4 * ndf), nn.LeakyReLU(0.2, True),
            nn.Conv2d(4 * ndf, 1, kernel_size=4, stride=1, padding=0, bias=True))

    def _build_exec(self, arena):
        m, ndf, lr = self.model, self.ndf, L.ACT_LRELU
        ls = [Layer("model.0", 0, m[0], self.input_nc, ndf, 3, 2, 1, act=lr),
              Layer("model.2", 1, m[2], ndf, 2 * ndf, 3, 2, 1, norm=_norm_kind(m[3]), act=lr, norm_mod=m[3]),
              Layer("model.5", 2, m[5], 2 * ndf, 4 * ndf, 3, 2, 1, norm=_norm_kind(m[6]), act=lr, norm_mod=m[6]),
              Layer("model.8", 3, m[8], 4 * ndf, 4 * ndf, 3, 2, 1, norm=_norm_kind(m[9]), act=lr, norm_mod=m[9]),
              Layer("out", 4, m[11], 4 * ndf, 1, 4, 1, 0, head=True)]
        return NetExec(self, ls, self.input_nc, 0, arena)

    def forward(self, input):
        return self._call(("out",), input)


class DiscriminatorLatent(_FusedNet):
    """networks.py:396-433: Linear/BatchNorm1d/LeakyReLU MLP, run as 1x1 convs on [N,1,1,C] planes."""

    def __init__(self, nlatent, ndf, use_sigmoid=False, gpu_ids=[]):
        super().__init__()
        if use_sigmoid:
            raise NotImplementedError("dtg_b200: LSGAN only")
        self.gpu_ids, self.nlatent, self.ndf = gpu_ids, nlatent, ndf
        self.model = nn.Sequential(
            nn.Linear(nlatent, ndf), nn.BatchNorm1d(ndf), nn.LeakyReLU(0.2, True),
            nn.Linear(ndf, ndf), nn.BatchNorm1d(ndf), nn.LeakyReLU(0.2, True),
            nn.Linear(ndf, ndf), nn.BatchNorm1d(ndf), nn.LeakyReLU(0.2, True),
            nn.Linear(ndf, 1))

    def _build_exec(self, arena):
        m, ndf, lr = self.model, self.ndf, L.ACT_LRELU
        ls = [Layer("model.0", 0, m[0], self.nlatent, ndf, 1, norm=_norm_kind(m[1]), act=lr, norm_mod=m[1]),
              Layer("model.3", 1, m[3], ndf, ndf, 1, norm=_norm_kind(m[4]), act=lr, norm_mod=m[4]),
              Layer("model.6", 2, m[6], ndf, ndf, 1, norm=_norm_kind(m[7]), act=lr, norm_mod=m[7]),
              Layer("out", 3, m[9], ndf, 1, 1, head=True)]
        return NetExec(self, ls, self.nlatent, 0, arena)

    def forward(self, input):
        n = input.size(0)
        out = self._call(("out",), input.reshape(n, self.nlatent, 1, 1))
        return out.reshape(n, 1)


class LatentEncoder(_FusedNet):
    """networks.py:438-482; returns (mu, logvar) flattened to [N, -1].

    The reference's encoder only yields [N, nlatent] codes for 64x64 inputs (four stride-2 convs then a 4x4 valid conv,
    SURVEY section 0): at 128x128 it returns [N, 25 * nlatent] and AugmentedCycleGAN fails.  img_size = 64 * 2^k is the
    N3 extension: k further stride-2 stages (conv 8nef -> 8nef, k3 s2 p1, no bias + norm + ReLU) are inserted
    before the 4x4 valid conv, so the code stays [N, nlatent]; img_size=64 builds exactly the reference's modules and
    state-dict keys."""

    def __init__(self, nlatent, input_nc, nef, norm_layer, gpu_ids=[], img_size=64):
        super().__init__()
        self.gpu_ids, self.nlatent, self.input_nc, self.nef = gpu_ids, nlatent, input_nc, nef
        extra, s = 0, int(img_size)
        while s > 64 and s % 2 == 0:
            s //= 2
            extra += 1
        if s != 64:
            raise ValueError("dtg_b200: LatentEncoder img_size must be 64 * 2^k, got %r" % (img_size,))
        self.img_size, self.n_extra = int(img_size), extra
        kw = 3
        seq = [
            nn.Conv2d(input_nc, nef, kernel_size=kw, stride=2, padding=1, bias=True), nn.ReLU(True),
            nn.Conv2d(nef, 2 * nef, kernel_size=kw, stride=2, padding=1, bias=False), norm_layer(2 * nef), nn.ReLU(True),
            nn.Conv2d(2 * nef, 4 * nef, kernel_size=kw, stride=2, padding=1, bias=False), norm_layer(4 * nef), nn.ReLU(True),
            nn.Conv2d(4 * nef, 8 * nef, kernel_size=kw, stride=2, padding=1, bias=False), norm_layer(8 * nef), nn.ReLU(True)]
        for _ in range(extra):
            seq += [nn.Conv2d(8 * nef, 8 * nef, kernel_size=kw, stride=2, padding=1, bias=False), norm_layer(8 * nef),
                    nn.ReLU(True)]
        seq += [nn.Conv2d(8 * nef, 8 * nef, kernel_size=4, stride=1, padding=0, bias=False), norm_layer(8 * nef), nn.ReLU(True)]
        self.conv_modules = nn.Sequential(*seq)
        self.enc_mu = nn.Conv2d(8 * nef, nlatent, kernel_size=1, stride=1, padding=0, bias=True)
        self.enc_logvar = nn.Conv2d(8 * nef, nlatent, kernel_size=1, stride=1, padding=0, bias=True)

    def _build_exec(self, arena):
        m, nef, r = self.conv_modules, self.nef, L.ACT_RELU
        ls = [Layer("conv_modules.0", 0, m[0], self.input_nc, nef, 3, 2, 1, act=r)]
        chans = [nef, 2 * nef, 4 * nef, 8 * nef] + [8 * nef] * self.n_extra
        for j in range(1, len(chans)):
            ci = 2 + 3 * (j - 1)
            ls.append(Layer("conv_modules.%d" % ci, j, m[ci], chans[j - 1], chans[j], 3, 2, 1, norm=_norm_kind(m[ci + 1]),
                            act=r, norm_mod=m[ci + 1]))
        ci, j = 2 + 3 * (len(chans) - 1), len(chans)
        ls += [Layer("conv_modules.%d" % ci, j, m[ci], 8 * nef, 8 * nef, 4, 1, 0, norm=_norm_kind(m[ci + 1]), act=r,
                     norm_mod=m[ci + 1]),
               Layer("mu", j + 1, self.enc_mu, 8 * nef, self.nlatent, 1, head=True),
               Layer("logvar", j + 1, self.enc_logvar, 8 * nef, self.nlatent, 1, head=True)]
        return NetExec(self, ls, self.input_nc, 0, arena)

    def forward(self, input):
        mu, logvar = self._call(("mu", "logvar"), input)
        return mu.reshape(mu.size(0), -1), logvar.reshape(logvar.size(0), -1)


# ---- factories (networks.py:33-127) ----------------------------------------------------------------

def define_G(input_nc, output_nc, ngf, norm='instance', which_model_netG='resnet', use_dropout=False, gpu_ids=[],
             n_blocks=None):
    """n_blocks=None: the reference's call (n_blocks=9 passed and ignored -> 3 blocks); an int is honoured (N3)"""
    if len(gpu_ids) > 0:
        assert (torch.cuda.is_available())
    netG = ResnetGenerator(input_nc, output_nc, ngf, norm_layer=get_norm_layer(norm_type=norm),
                           use_dropout=use_dropout, n_blocks=9 if n_blocks is None else n_blocks, gpu_ids=gpu_ids,
                           honor_n_blocks=n_blocks is not None)
    if len(gpu_ids) > 0:
        netG.cuda()
    netG.apply(weights_init)
    return netG


def define_stochastic_G(nlatent, input_nc, output_nc, ngf, norm='instance', which_model_netG='resnet',
                        use_dropout=False, gpu_ids=[], n_blocks=None):
    if len(gpu_ids) > 0:
        assert (torch.cuda.is_available())
    netG = CINResnetGenerator(nlatent, input_nc, output_nc, ngf, norm_layer=CondInstanceNorm,
                              use_dropout=use_dropout, n_blocks=9 if n_blocks is None else n_blocks, gpu_ids=gpu_ids,
                              honor_n_blocks=n_blocks is not None)
    if len(gpu_ids) > 0:
        netG.cuda()
    netG.apply(weights_init)
    return netG


def define_D_A(input_nc, ndf, which_model_netD, norm, use_sigmoid=False, gpu_ids=[]):
    if len(gpu_ids) > 0:
        assert (torch.cuda.is_available())
    netD = Discriminator_edges(input_nc, ndf, norm_layer=get_norm_layer(norm_type=norm), use_sigmoid=use_sigmoid,
                               gpu_ids=gpu_ids)
    if len(gpu_ids) > 0:
        netD.cuda()
    netD.apply(weights_init)
    return netD


def define_D_B(input_nc, ndf, which_model_netD, norm, use_sigmoid=False, gpu_ids=[]):
    if len(gpu_ids) > 0:
        assert (torch.cuda.is_available())
    netD = Discriminator(input_nc, ndf, norm_layer=get_norm_layer(norm_type=norm), use_sigmoid=use_sigmoid,
                         gpu_ids=gpu_ids)
    if len(gpu_ids) > 0:
        netD.cuda()
    netD.apply(weights_init)
    return netD


def define_LAT_D(nlatent, ndf, use_sigmoid=False, gpu_ids=[]):
    if len(gpu_ids) > 0:
        assert (torch.cuda.is_available())
    netD = DiscriminatorLatent(nlatent, ndf, use_sigmoid=use_sigmoid, gpu_ids=gpu_ids)
    if len(gpu_ids) > 0:
        netD.cuda()
    netD.apply(weights_init)
    return netD


def define_E(nlatent, input_nc, nef, norm='batch', gpu_ids=[], img_size=64):
    if len(gpu_ids) > 0:
        assert (torch.cuda.is_available())
    netE = LatentEncoder(nlatent, input_nc, nef, norm_layer=get_norm_layer(norm_type=norm), gpu_ids=gpu_ids,
                         img_size=img_size)
    if len(gpu_ids) > 0:
        netE.cuda()
    netE.apply(weights_init)
    return netE


def print_network(net, out_f=None):
    """networks.py:130-138"""
    num_params = sum(p.numel() for p in net.parameters())
    if out_f is not None:
        out_f.write(net.__repr__() + "\n")
        out_f.write('Total number of parameters: %d\n' % num_params)
        out_f.flush()
