"""Building-block modules with the reference's names, constructor signatures and state_dict keys
(/root/reference/augmented_cyclegan/modules.py:15-235).

Inside the networks these modules are PARAMETER HOLDERS: the network-level engine (engine.py) reads
their parameters and runs fused sm_100a kernels; it never calls their ``forward``.  Called on their own,
``InstanceNorm`` / ``CondInstanceNorm`` run the same fused norm kernels through an autograd Function
(fp32 NCHW in / out, like the reference), and ``CINResnetBlock`` / ``ResnetBlock`` run a two-layer fused plan
(reflect-padded 3x3 tcgen05 convs + fused norms + residual), so they remain drop-in usable as standalone layers.
"""
import weakref

import torch
import torch.nn as nn
from torch.nn.parameter import Parameter

from . import _lib as L
from . import ops


class TwoInputModule(nn.Module):
    """modules.py:15-17"""

    def forward(self, input1, input2):
        raise NotImplementedError


class MergeModule(TwoInputModule):
    """modules.py:25-37: o = module2(module1(x), z)"""

    def __init__(self, module1, module2):
        super().__init__()
        self.module1 = module1
        self.module2 = module2

    def forward(self, input1, input2):
        return self.module2.forward(self.module1.forward(input1), input2)


class TwoInputSequential(nn.Sequential, TwoInputModule):
    """modules.py:44-56"""

    def forward(self, input1, input2):
        for module in self._modules.values():
            if isinstance(module, TwoInputModule):
                input1 = module.forward(input1, input2)
            else:
                input1 = module.forward(input1)
        return input1


class _NormFn(torch.autograd.Function):
    """Standalone fused norm (+affine) on fp32 NCHW tensors via the fp32-plane kernels."""

    @staticmethod
    def forward(ctx, x, gamma, beta, mode, eps):
        n, c, h, w = x.shape
        xp = ops.PlaneT.from_nchw(x.detach().contiguous().float(), dtype=torch.float32)
        out = ops.PlaneT(n, h, w, xp.c, 0, torch.float32, x.device)
        st = ops.NormState(xp)
        cs = xp.c
        if mode == L.NORM_COND_INSTANCE:
            g = torch.zeros(n, cs, device=x.device); g[:, :c] = gamma.detach().reshape(n, c)
            b = torch.zeros(n, cs, device=x.device); b[:, :c] = beta.detach().reshape(n, c)
        else:
            g = torch.zeros(cs, device=x.device); g[:c] = gamma.detach()
            b = torch.zeros(cs, device=x.device); b[:c] = beta.detach()
        ops.norm_fwd(xp, out, st, mode=mode, act=L.ACT_NONE, gamma=g, beta=b, eps=eps)
        ctx.saved = (xp, st, g, mode, c)
        ctx.gshape = gamma.shape
        return ops.unpack_nchw(out, c)

    @staticmethod
    def backward(ctx, dy):
        xp, st, g, mode, c = ctx.saved
        dyp = ops.PlaneT.from_nchw(dy.contiguous().float(), dtype=torch.float32, c_store=xp.c)
        dx = ops.PlaneT(xp.n, xp.h, xp.w, xp.c, 0, torch.float32, dy.device)
        dg = torch.zeros(xp.c, device=dy.device)
        db = torch.zeros(xp.c, device=dy.device)
        ops.norm_bwd(dyp, dx, st, mode=mode, act=L.ACT_NONE, x=xp, gamma=g, d_gamma=dg, d_beta=db, want_sums=True)
        if mode == L.NORM_COND_INSTANCE:
            dgam = st.sums[:, :c, 1].reshape(ctx.gshape).clone()
            dbet = st.sums[:, :c, 0].reshape(ctx.gshape).clone()
        else:
            dgam, dbet = dg[:c], db[:c]
        return ops.unpack_nchw(dx, c), dgam, dbet, None, None


class InstanceNorm(nn.Module):
    """modules.py:64-97: per-(n,c) mean / BIASED variance, per-channel scale ~ N(0, .02), shift 0."""

    def __init__(self, num_features, affine=True, eps=1e-5):
        super().__init__()
        self.num_features = num_features
        self.affine = affine
        self.eps = eps
        self.scale = Parameter(torch.Tensor(num_features))
        self.shift = Parameter(torch.Tensor(num_features))
        self.reset_parameters()

    def reset_parameters(self):
        if self.affine:
            self.scale.data.normal_(mean=0., std=0.02)
            self.shift.data.zero_()

    def forward(self, input):
        if self.affine:
            return _NormFn.apply(input, self.scale, self.shift, L.NORM_INSTANCE, self.eps)
        one = torch.ones(self.num_features, device=input.device)
        return _NormFn.apply(input, one, torch.zeros_like(one), L.NORM_INSTANCE, self.eps)


InstanceNorm2d = InstanceNorm   # modules.py:98 (USE_PYTORCH_IN = False)


class _CinAffineFn(torch.autograd.Function):
    """gamma = relu(Ws z + bs), beta = relu(Wb z + bb) through dtg_cin_affine_{fwd,bwd}."""

    @staticmethod
    def forward(ctx, z, ws, bs, wb, bb):
        n, nz = z.shape[0], z.shape[1]
        c = ws.shape[0]
        z2 = z.detach().reshape(n, nz).contiguous().float()
        gam = torch.empty(n, c, device=z.device)
        bet = torch.empty(n, c, device=z.device)
        ops.cin_affine_fwd(z2, ws.detach().contiguous(), bs.detach().contiguous(), wb.detach().contiguous(),
                           bb.detach().contiguous(), gam, bet)
        ctx.saved = (z2, ws.detach(), wb.detach(), gam, bet, z.shape)
        return gam, bet

    @staticmethod
    def backward(ctx, dgam, dbet):
        z2, ws, wb, gam, bet, zshape = ctx.saved
        sums = torch.stack([dbet, dgam], dim=-1).contiguous()   # (sum g, sum g*xhat) = (d_shift, d_scale)
        d = [torch.zeros_like(ws), torch.zeros(ws.shape[0], device=ws.device), torch.zeros_like(wb),
             torch.zeros(wb.shape[0], device=wb.device), torch.zeros_like(z2)]
        ops.cin_affine_bwd(z2, ws.contiguous(), wb.contiguous(), gam, bet, sums, *d)
        return d[4].reshape(zshape), d[0], d[1], d[2], d[3]


class CondInstanceNorm(TwoInputModule):
    """modules.py:104-132: z-projected non-negative scale / shift, UNBIASED variance."""

    def __init__(self, x_dim, z_dim, eps=1e-5):
        super().__init__()
        self.eps = eps
        self.shift_conv = nn.Sequential(nn.Conv2d(z_dim, x_dim, kernel_size=1, padding=0, bias=True), nn.ReLU(True))
        self.scale_conv = nn.Sequential(nn.Conv2d(z_dim, x_dim, kernel_size=1, padding=0, bias=True), nn.ReLU(True))

    def forward(self, input, noise):
        sc, sh = self.scale_conv[0], self.shift_conv[0]
        gam, bet = _CinAffineFn.apply(noise, sc.weight, sc.bias, sh.weight, sh.bias)
        return _NormFn.apply(input, gam, bet, L.NORM_COND_INSTANCE, self.eps)


def _release_ctx(c):
    c.busy = False


class _BlockFn(torch.autograd.Function):
    """Standalone residual block (modules.py:185-188, 232-235): out = relu(x + conv_block(x[, z])) as a two-layer
    engine plan on a private context; differentiable in x, z and the block's parameters (gradients accumulate into
    ``.grad``)."""

    @staticmethod
    def forward(ctx, block, x, z):
        ex = block._exec()
        n, cch, h, w = x.shape
        ctx.set_materialize_grads(False)
        ex.repack()
        slot = 0
        while getattr(ex.new_ctx(n, h, w, tag=("autograd", slot)), "busy", False):
            slot += 1
        c = ex.new_ctx(n, h, w, tag=("autograd", slot))
        c.busy = torch.is_grad_enabled()
        if c.busy:
            weakref.finalize(ctx, _release_ctx, c)
        ops.pack_nchw(x.detach().contiguous().float(), c.acts[0], 0)        # mirrors into the halo: ReflectionPad2d(1)
        if z is not None:
            c.z.copy_(z.detach().reshape(n, -1))
        ex.forward(c)
        ctx.block, ctx.c, ctx.cch = block, c, cch
        ctx.zshape = z.shape if z is not None else None
        return ops.unpack_nchw(c.acts[len(ex.layers)], cch)

    @staticmethod
    def backward(ctx, dy):
        if dy is None:
            return None, None, None
        block, c = ctx.block, ctx.c
        ex = block._exec()
        c.busy = False
        top = c.acts[len(ex.layers)]
        if not hasattr(c, "top_grad"):
            c.top_grad = ops.PlaneT(top.n, top.h, top.w, top.c, 0, top.dtype, top.t.device)
        ops.pack_nchw(dy.contiguous().float(), c.top_grad, 0, reflect=False)
        gin, dres = ex.backward(c, {}, want_dx="pair", want_dw=True, want_dz=ctx.zshape is not None, top_grad=c.top_grad)
        dx = torch.empty(c.n, ctx.cch, gin.h, gin.w, device=gin.t.device)
        ops.grad_gather([gin, dres], [0, 0], ctx.cch, out=None, out_nchw=dx)     # dgrad of conv 1 (ring folded) + identity branch
        dz = c.dz.reshape(ctx.zshape).clone() if ctx.zshape is not None else None
        return None, dx, dz


class _FusedBlock(object):
    """lazy two-layer plan of a standalone residual block (engine.LooseArena: the block does not own its weights)"""

    def _exec(self):
        from . import engine
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("dtg_b200 modules run only on CUDA (sm_100a); there is no CPU fallback")
        ex = getattr(self, "_ex", None)
        if ex is None or ex.arena.device != dev:
            ex = self._build_exec(engine)
            object.__setattr__(self, "_ex", ex)
        ex.prepare()
        return ex


def _pad_layers(padding_type):
    if padding_type == 'reflect':
        return [nn.ReflectionPad2d(1)], 0
    if padding_type == 'replicate':
        return [nn.ReplicationPad2d(1)], 0
    if padding_type == 'zero':
        return [], 1
    raise NotImplementedError('padding [%s] is not implemented' % padding_type)


class CINResnetBlock(TwoInputModule, _FusedBlock):
    """modules.py:139-188 (same child indices and the `str(idx)` aliases of :145-146).  Inside CINResnetGenerator the
    block is a parameter holder of the generator's plan; called on its own it runs the two-layer plan below."""

    def __init__(self, x_dim, z_dim, padding_type, norm_layer, use_dropout, use_bias):
        super().__init__()
        if use_dropout or padding_type != 'reflect':
            raise NotImplementedError("dtg_b200: residual blocks implement reflect padding without dropout "
                                      "(the reference's configuration, networks.py:173-175)")
        self.x_dim, self.z_dim = x_dim, z_dim
        self.padding_type = padding_type
        self.conv_block = self.build_conv_block(x_dim, z_dim, padding_type, norm_layer, use_dropout, use_bias)
        self.relu = nn.ReLU(True)
        for idx, module in enumerate(self.conv_block):
            self.add_module(str(idx), module)

    def build_conv_block(self, x_dim, z_dim, padding_type, norm_layer, use_dropout, use_bias):
        conv_block = []
        pads, p = _pad_layers(padding_type)
        conv_block += pads
        conv_block += [MergeModule(nn.Conv2d(x_dim, x_dim, kernel_size=3, padding=p, bias=use_bias),
                                   norm_layer(x_dim, z_dim)), nn.ReLU(True)]
        if use_dropout:
            conv_block += [nn.Dropout(0.5)]
        pads, p = _pad_layers(padding_type)
        conv_block += pads
        conv_block += [nn.Conv2d(x_dim, x_dim, kernel_size=3, padding=p, bias=use_bias),
                       InstanceNorm2d(x_dim, affine=True)]
        return TwoInputSequential(*conv_block)

    def _build_exec(self, engine):
        cb, d = self.conv_block, self.x_dim
        conv1, n1 = cb[1].module1, cb[1].module2
        kind = L.NORM_COND_INSTANCE if isinstance(n1, CondInstanceNorm) else L.NORM_INSTANCE
        ls = [engine.Layer("a", 0, conv1, d, d, 3, 1, 1, norm=kind, act=L.ACT_RELU, norm_mod=n1, out_halo=1),
              engine.Layer("b", 1, cb[4], d, d, 3, 1, 1, norm=L.NORM_INSTANCE, act=L.ACT_RELU, norm_mod=cb[5], residual=0)]
        return engine.NetExec(self, ls, d, 1, engine.LooseArena(self), nz=self.z_dim)

    def forward(self, x, noise):
        """modules.py:185-188: relu(x + conv_block(x, noise))"""
        return _BlockFn.apply(self, x, noise)


class ResnetBlock(nn.Module, _FusedBlock):
    """modules.py:193-235.  Inside ResnetGenerator a parameter holder; standalone it runs the two-layer plan below."""

    def __init__(self, dim, padding_type, norm_layer, use_dropout, use_bias):
        super().__init__()
        if use_dropout or padding_type != 'reflect':
            raise NotImplementedError("dtg_b200: residual blocks implement reflect padding without dropout "
                                      "(the reference's configuration, networks.py:225-227)")
        self.dim = dim
        self.padding_type = padding_type
        self.conv_block = self.build_conv_block(dim, padding_type, norm_layer, use_dropout, use_bias)
        self.relu = nn.ReLU(True)

    def build_conv_block(self, dim, padding_type, norm_layer, use_dropout, use_bias):
        conv_block = []
        pads, p = _pad_layers(padding_type)
        conv_block += pads
        conv_block += [nn.Conv2d(dim, dim, kernel_size=3, padding=p, bias=use_bias), nn.ReLU(True)]
        if use_dropout:
            conv_block += [nn.Dropout(0.5)]
        pads, p = _pad_layers(padding_type)
        conv_block += pads
        conv_block += [nn.Conv2d(dim, dim, kernel_size=3, padding=p, bias=use_bias), norm_layer(dim)]
        return nn.Sequential(*conv_block)

    def _build_exec(self, engine):
        cb, d = self.conv_block, self.dim
        if not isinstance(cb[5], InstanceNorm):
            raise NotImplementedError("dtg_b200: standalone ResnetBlock implements the reference's InstanceNorm2d")
        ls = [engine.Layer("a", 0, cb[1], d, d, 3, 1, 1, norm=L.NORM_NONE, act=L.ACT_RELU, out_halo=1),
              engine.Layer("b", 1, cb[4], d, d, 3, 1, 1, norm=L.NORM_INSTANCE, act=L.ACT_RELU, norm_mod=cb[5], residual=0)]
        return engine.NetExec(self, ls, d, 1, engine.LooseArena(self))

    def forward(self, x):
        """modules.py:232-235: relu(x + conv_block(x))"""
        return _BlockFn.apply(self, x, None)
