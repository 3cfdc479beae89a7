"""Execution engine: flat parameter arenas, per-network layer plans, forward/backward drivers.

A network (networks.py) is described as a chain of ``Layer`` specs (conv [+ norm + activation
(+ residual)]).  ``NetExec`` owns the packed tensor-core operands of one network and runs a plan on a
``Ctx`` (the saved activations of ONE invocation) purely through the C ABI (ops.py): no torch
arithmetic happens here.  Gradients are accumulated straight into the flat fp32 gradient arena in
PyTorch parameter layout, so ``param.grad`` views and the fused clip+Adam kernel see them in place.
"""
import os

import functools

import torch

from . import _lib as L
from . import ops

_PRECISION = {"dtype": torch.bfloat16}
S2D = os.environ.get("DTG_NO_S2D") is None      # space-to-depth execution of the stride-2 input layers
# matrix-vector kernels (dtg_head1_*) for the single-output-channel PatchGAN heads: measured equal to the N-padded
# tensor-core path (fwd 52 vs 44 us, dgrad 49 vs 55 us, wgrad 68 vs 59 us at batch 160; step time unchanged), so the
# tensor-core path stays the default and this is opt-in
HEAD1 = os.environ.get("DTG_HEAD1") is not None
FLAT_DGRAD = os.environ.get("DTG_NO_FLAT_DGRAD") is None      # haloed dy planes for the reflect-ring data gradients (conv_patch2.cu flat mode)
OFFCHAIN_SMALL = os.environ.get("DTG_NO_OFFCHAIN_SMALL") is None      # parameter-gradient reductions of the norm layers on the companion stream
TAIL_KWN = os.environ.get("DTG_NO_TAIL_KWN") is None      # generators' 7x7 tail forward with (kw, cout) in GEMM-N (conv_tail7.cu)


def set_precision(name):
    """'bf16' (tcgen05 kind::f16, bf16 planes) or 'tf32' (kind::tf32, fp32 planes)."""
    _PRECISION["dtype"] = {"bf16": torch.bfloat16, "tf32": torch.float32}[name]


def get_dtype():
    return _PRECISION["dtype"]


def running_pair(bn):
    """running_mean / running_var of a BatchNorm module as rows of ONE [2, c] fp32 buffer (what
    dtg_norm_fwd updates in place); the module's registered buffers become views of it, so
    state_dict() / load_state_dict() keep working."""
    rp = getattr(bn, "_dtg_pair", None)
    if rp is None or rp.device != bn.running_mean.device or bn.running_mean.data_ptr() != rp.data_ptr():
        rp = torch.stack([bn.running_mean.detach().float(), bn.running_var.detach().float()]).contiguous()
        bn.running_mean = rp[0]
        bn.running_var = rp[1]
        object.__setattr__(bn, "_dtg_pair", rp)
    return rp


class ParamArena:
    """All parameters of one network in one contiguous fp32 buffer (+ grad, Adam m / v).

    Each nn.Parameter's ``.data`` becomes a view into ``flat`` (state_dict / load_state_dict keep
    working in place) and ``.grad`` a view into ``grad``.  Parameters listed in `inactive` (no gradient
    in the reference, e.g. enc_logvar with stoch_enc=False) are placed last and excluded from
    clip / Adam, like torch skips ``grad is None`` parameters.
    """

    def __init__(self, module, inactive=()):
        seen, plist = set(), []
        for name, p in module.named_parameters():   # named_parameters de-duplicates aliases
            if id(p) not in seen:
                seen.add(id(p))
                plist.append((name, p))
        act = [(n, p) for n, p in plist if n not in inactive]
        ina = [(n, p) for n, p in plist if n in inactive]
        self.module = module
        self.entries = act + ina
        self.inactive_names = set(inactive)
        self.flat = None
        self._build()

    def _build(self):
        dev = self.entries[0][1].device
        total, offs = 0, []
        for _, p in self.entries:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4          # keep every tensor 16-byte aligned
        self.active_count = 0
        for (n, p), o in zip(self.entries, offs):
            if n not in self.inactive_names:
                self.active_count = o + (p.numel() + 3) // 4 * 4
        self.total = total
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.m = torch.zeros(total, dtype=torch.float32, device=dev)
        self.v = torch.zeros(total, dtype=torch.float32, device=dev)
        self.views, self.gviews = {}, {}
        for (n, p), o in zip(self.entries, offs):
            v = self.flat[o:o + p.numel()].view(p.shape)
            v.copy_(p.data)
            p.data = v
            g = self.grad[o:o + p.numel()].view(p.shape)
            p.grad = None if n in self.inactive_names else g
            self.views[n] = v
            self.gviews[n] = g
        self.device = dev

    def ensure(self):
        """Re-flatten if the module was moved (.cuda()/.to()) after construction."""
        p0 = self.entries[0][1]
        if p0.device != self.device or p0.data_ptr() != self.flat.data_ptr():
            self._build()
            return True
        for (n, p) in self.entries:   # .grad may have been reset by zero_grad(set_to_none=True)
            if n not in self.inactive_names and (p.grad is None or p.grad.data_ptr() != self.gviews[n].data_ptr()):
                p.grad = self.gviews[n]
        return False

    def g(self, param):
        """gradient view of a parameter object"""
        for (n, p) in self.entries:
            if p is param:
                return self.gviews[n]
        raise KeyError("parameter not in arena")


class LooseArena:
    """Arena facade for a module that runs a fused plan on its own but does NOT own its parameters' storage (a
    CINResnetBlock / ResnetBlock called standalone may live inside a generator whose ParamArena holds the weights):
    nothing is re-pointed; gradients accumulate straight into each parameter's ``.grad`` (allocated on demand)."""

    def __init__(self, module):
        self.module = module
        self.device = next(module.parameters()).device
        self._ptrs = None

    def ensure(self):
        self.device = next(self.module.parameters()).device
        ptrs = tuple(p.data_ptr() for p in self.module.parameters())
        changed = ptrs != self._ptrs          # weights moved (an outer arena re-flattened them, .cuda(), ...): repack
        self._ptrs = ptrs
        return changed

    def g(self, param):
        if param.grad is None or not param.grad.is_contiguous() or param.grad.dtype != torch.float32:
            param.grad = torch.zeros_like(param, dtype=torch.float32, memory_format=torch.contiguous_format)
        return param.grad


class Layer:
    """conv (+ norm + act (+ residual)) -> activation plane, or a head conv -> dense NCHW fp32."""

    def __init__(self, name, src, conv, cin, cout, k, stride=1, pad=0, transposed=False, norm=L.NORM_NONE,
                 act=L.ACT_NONE, norm_mod=None, out_halo=0, residual=None, head=False, use_bias=True,
                 spatial=None):
        self.name, self.src, self.conv = name, src, conv
        self.cin, self.cout, self.k, self.stride, self.pad, self.transposed = cin, cout, k, stride, pad, transposed
        self.norm, self.act, self.norm_mod = norm, act, norm_mod
        self.out_halo, self.residual, self.head = out_halo, residual, head
        # a bias in front of a mean-removing norm is a mathematical no-op: skipped (its gradient is exactly
        # 0).  Batch norm keeps it in the forward because running_mean must include it.
        self.use_bias = use_bias and norm == L.NORM_NONE and conv.bias is not None
        self.fwd_bias = self.use_bias or (norm == L.NORM_BATCH and conv.bias is not None)
        self.w_f = self.w_d = self.w_f_s2d = None
        # kw-folded execution of the small-channel 7x7 layers (decided in NetExec.prepare for the plane dtype):
        # fold_in : the INPUT is a 16-byte-per-pixel plane  -> forward conv and wgrad read it kw-folded
        # fold_out: the OUTPUT gradient is (head with <= fc channels) -> dgrad and wgrad read dy kw-folded
        self.fold_in = self.fold_out = False
        self.tail_kwn, self.w_f_kwn = False, None
        self.head_kwn_d, self.w_d_kwn = False, None
        # space-to-depth execution of a stride-2 first layer with a <= 64-byte input pixel (decided in
        # NetExec.prepare): the input plane is stored as 2x2 pixel blocks, forward conv and wgrad run as stride-1 3x3
        self.s2d = False
        # single-output-channel head (PatchGAN last layer): matrix-vector shaped, runs on the dtg_head1_* kernels
        self.head1 = False

    def out_hw(self, h, w):
        if self.transposed:       # k3 s2 p1 op1 -> exactly 2x
            return h * self.stride, w * self.stride
        return (h + 2 * self.pad - self.k) // self.stride + 1, (w + 2 * self.pad - self.k) // self.stride + 1


class Ctx:
    """Saved state of one forward invocation of a network (static buffers, CUDA-graph friendly)."""

    def __init__(self):
        self.acts = {}      # index -> PlaneT (0 = packed input)
        self.yraw = {}      # layer idx -> PlaneT (conv output before the norm)
        self.nst = {}       # layer idx -> NormState
        self.cin = {}       # layer idx -> (gamma, beta) [n, c]
        self.gact = {}      # act index -> gradient plane (ring = act halo)
        self.dres = {}      # act index -> residual-branch gradient plane
        self.dyraw = {}     # layer idx -> gradient w.r.t. the conv output
        self.heads = {}     # layer name -> dense NCHW fp32 output
        self.z = None       # [n, nz] fp32 (CIN networks)
        self.dz = None


class NetExec:
    def __init__(self, module, layers, in_channels, in_halo, arena, nz=0):
        self.module, self.layers, self.arena = module, layers, arena
        self.in_channels, self.in_halo, self.nz = in_channels, in_halo, nz
        self.dtype = None
        self.pack = None
        self._ctx_cache = {}

    # ---- operands ----------------------------------------------------------------------------
    def prepare(self, dtype=None):
        dtype = dtype or get_dtype()
        rebuilt = self.arena.ensure()
        if self.pack is None or rebuilt or dtype != self.dtype:
            self.dtype = dtype
            self.pack = ops.PackTable(self.arena.device)
            fc = ops.fold_channels(dtype)
            cp = ops.cpad_small(self.in_channels, dtype)
            first = [ly for ly in self.layers if ly.src == 0]
            s2d_ok = (S2D and self.in_halo == 0 and 4 * cp * (2 if dtype == torch.bfloat16 else 4) <= 256 and
                      all((not ly.transposed) and ly.stride == 2 and ly.pad == 1 and ly.k in (3, 4) and
                          ly.conv.weight.dim() == 4 for ly in first))
            self.s2d_cp = cp if (s2d_ok and first) else 0
            for ly in self.layers:
                ly.s2d = bool(self.s2d_cp) and ly.src == 0
                w = ly.conv.weight
                w4 = w if w.dim() == 4 else w.view(w.shape[0], w.shape[1], 1, 1)
                foldable = (not ly.transposed) and ly.stride == 1 and 3 < ly.k <= 8 and 2 * ly.pad == ly.k - 1
                ly.fold_in = foldable and ly.src == 0 and ly.cin <= fc and self.in_channels <= fc and self.in_halo >= ly.pad
                src_halo = self.in_halo if ly.src == 0 else self.layers[ly.src - 1].out_halo
                ly.fold_out = foldable and ly.head and ly.cout <= fc and not ly.fold_in and src_halo == 0
                # 7x7 'same' head with <= 4 output channels on a halo-free 32/64/128-byte-per-pixel plane: forward with the
                # filter column in GEMM-N (conv_tail7.cu) wherever the image width divides 128 (decided per context)
                rowb = ops.cpad(ly.cin, dtype) * (2 if dtype == torch.bfloat16 else 4)
                ly.tail_kwn = (TAIL_KWN and foldable and ly.head and ly.cout <= 4 and ly.k * ly.cout <= 28 and src_halo == 0
                               and ly.src > 0 and rowb in (32, 64, 128) and ops.cpad(ly.cin, dtype) == ly.cin)
                ly.w_f_kwn = ops.add_packed(self.pack, w4, dtype, "fwd_kwn") if ly.tail_kwn else None
                # the mirror image: data gradient of the 7x7 HEAD (16-byte-pixel input with a materialised halo == pad)
                ly.head_kwn_d = (TAIL_KWN and bool(ly.fold_in) and ly.cin <= 4 and ly.k * ly.cin <= 28 and self.in_halo == ly.pad
                                 and ops.cpad(ly.cout, dtype) == ly.cout)
                ly.w_d_kwn = ops.add_packed(self.pack, w4, dtype, "dgrad_kwn") if ly.head_kwn_d else None
                ly.head1 = (HEAD1 and ly.head and ly.cout == 1 and not ly.transposed and ly.stride == 1 and w.dim() == 4 and
                            ly.k * ly.k <= 16 and ly.act == L.ACT_NONE and src_halo == 0 and ly.src > 0 and
                            ly.cin % 8 == 0 and ly.cin <= 512)
                if ly.head1:
                    ly.fold_in = ly.fold_out = False
                    continue            # no packed operand: the kernels read the fp32 master weight
                if ly.s2d:      # both forms: the block form needs even image extents (decided per context)
                    ly.w_f_s2d = ops.add_packed(self.pack, w4, dtype, "fwd_s2d", s2d_cp=cp)
                    ly.w_f = ops.add_packed(self.pack, w4, dtype, "fwd")
                    ly.w_d = ops.add_packed(self.pack, w4, dtype, "dgrad")
                elif ly.fold_in:
                    ly.w_f = ops.add_packed(self.pack, w4, dtype, "fwd_fold")
                    ly.w_d = ops.add_packed(self.pack, w4, dtype, "dgrad")
                elif ly.fold_out:
                    ly.w_f = ops.add_packed(self.pack, w4, dtype, "fwd")
                    ly.w_d = ops.add_packed(self.pack, w4, dtype, "dgrad_fold")
                elif ly.transposed:
                    ly.w_f = ops.add_packed(self.pack, w4, dtype, "tfwd")
                    ly.w_d = ops.add_packed(self.pack, w4, dtype, "tdgrad")
                else:
                    ly.w_f = ops.add_packed(self.pack, w4, dtype, "fwd")
                    ly.w_d = ops.add_packed(self.pack, w4, dtype, "dgrad")
            self._ctx_cache = {}
            self.repack()

    def repack(self):
        """refresh the packed bf16/tf32 operands from the fp32 master weights (after an optimizer step)"""
        self.pack.run()

    # ---- contexts ----------------------------------------------------------------------------
    def new_ctx(self, n, h, w, tag=0):
        key = (n, h, w, tag)
        if key in self._ctx_cache:
            return self._ctx_cache[key]
        dt, dev = self.dtype, self.arena.device
        c = Ctx()
        c.n = n
        c.s2d_cp = self.s2d_cp if (h % 2 == 0 and w % 2 == 0) else 0      # odd extents: ordinary strided path
        if c.s2d_cp:
            c.acts[0] = ops.PlaneT(n, h // 2, w // 2, 4 * self.s2d_cp, 0, dt, dev, s2d=self.s2d_cp)
            c.s2d_dw2 = {i: torch.zeros(ly.cout * 4 * self.s2d_cp * 9, dtype=torch.float32, device=dev)
                         for i, ly in enumerate(self.layers) if ly.s2d}
        else:
            c.acts[0] = ops.PlaneT(n, h, w, ops.cpad_small(self.in_channels, dt), self.in_halo, dt, dev)
        dims = {0: (h, w)}
        for i, ly in enumerate(self.layers):
            ih, iw = dims[ly.src]
            oh, ow = ly.out_hw(ih, iw)
            cs = ops.cpad(ly.cout, dt)
            if ly.head:
                c.heads[ly.name] = torch.zeros(n, ly.cout, oh, ow, dtype=torch.float32, device=dev)
                if ly.fold_out:     # seed gradient plane read kw-folded: 16-byte pixels, zero halo = the conv's padding
                    c.dyraw[i] = ops.PlaneT(n, oh, ow, ops.fold_channels(dt), ly.pad, dt, dev)
                else:
                    c.dyraw[i] = ops.PlaneT(n, oh, ow, cs, 0, dt, dev)      # seed gradient plane
                continue
            dims[i + 1] = (oh, ow)
            c.acts[i + 1] = ops.PlaneT(n, oh, ow, cs, ly.out_halo, dt, dev)
            if ly.norm != L.NORM_NONE:
                c.yraw[i] = ops.PlaneT(n, oh, ow, cs, 0, dt, dev)
                c.nst[i] = ops.NormState(c.yraw[i])
                if ly.norm == L.NORM_COND_INSTANCE:
                    c.cin[i] = (torch.zeros(n, cs, device=dev), torch.zeros(n, cs, device=dev))
            else:
                c.nst[i] = ops.NormState(c.acts[i + 1])
            # gradient w.r.t. the conv output.  For a stride-1 'same' conv over a reflect-padded input (the residual stack)
            # it gets a ZERO halo ring = the ring of the data gradient: the flat-raster dgrad then reads the ring as padding
            # (722 instead of 1200 tiles at 80 x 32 x 32).  Its producer must be the register-resident norm backward
            # (<= 1024 pixels per slab), which leaves the ring untouched; the weight gradient views the interior.
            src_halo = c.acts[ly.src].halo
            flat = (FLAT_DGRAD and not ly.transposed and ly.stride == 1 and 2 * ly.pad == ly.k - 1 and ly.pad >= 1 and
                    src_halo == ly.pad and not (ly.src == 0 and c.s2d_cp) and oh * ow <= 1024 and
                    ly.norm in (L.NORM_NONE, L.NORM_INSTANCE, L.NORM_COND_INSTANCE) and
                    ops.flat_dgrad_eligible(n, oh, ow, cs, c.acts[ly.src].c, ly.pad, dt))
            c.dyraw[i] = ops.PlaneT(n, oh, ow, cs, ly.pad if flat else 0, dt, dev)
        c.dims = dims
        if self.nz:
            c.z = torch.zeros(n, self.nz, dtype=torch.float32, device=dev)
            c.dz = torch.zeros(n, self.nz, dtype=torch.float32, device=dev)
        self._ctx_cache[key] = c
        return c

    def _gact(self, c, idx, consumer):
        """gradient plane of activation `idx` produced by the dgrad of layer `consumer`"""
        key = (idx, consumer)
        if key not in c.gact:
            a = c.acts[idx]
            if idx == 0 and c.s2d_cp:         # the input gradient keeps the ordinary pixel layout
                h, w = c.dims[0]
                c.gact[key] = ops.PlaneT(a.n, h, w, self.s2d_cp, 0, a.dtype, a.t.device)
            else:
                c.gact[key] = ops.PlaneT(a.n, a.h, a.w, a.c, a.halo, a.dtype, a.t.device)
        return c.gact[key]

    def _dres(self, c, idx):
        if idx not in c.dres:
            a = c.acts[idx]
            c.dres[idx] = ops.PlaneT(a.n, a.h, a.w, a.c, 0, a.dtype, a.t.device)
        return c.dres[idx]

    # ---- forward -----------------------------------------------------------------------------
    def forward(self, c, sync_bn=None):
        """c.acts[0] (and c.z) must be filled.  Returns dict of head outputs."""
        for i, ly in enumerate(self.layers):
            a_in = c.acts[ly.src]
            ih, iw = c.dims[ly.src]
            oh, ow = ly.out_hw(ih, iw)
            mode = L.CONV_DGRAD if ly.transposed else L.CONV_FWD
            kw = dict(mode=mode, kh=ly.k, kw=ly.k, stride=ly.stride, pad=ly.pad, cout=ly.cout, out_h=oh, out_w=ow,
                      cin=ly.cin, fold_w=ly.fold_in)
            w_f = ly.w_f
            if ly.s2d and c.s2d_cp:      # stride-1 3x3 over 2x2 pixel blocks; cin keeps the algorithmic FLOP count of the KxK filter
                kw.update(kh=3, kw=3, stride=1, pad=1, cin=ly.cin * ly.k * ly.k / 9.0)
                w_f = ly.w_f_s2d
            bias = ly.conv.bias if ly.use_bias else None
            if ly.head1:
                ops.head1_fwd(a_in, ly.conv.weight, bias, c.heads[ly.name], ly.pad)
                continue
            if ly.head:
                if ly.tail_kwn and ops.tail_kwn_eligible(a_in.c, ly.k, ly.cout, iw, a_in.dtype):
                    kw.update(fold_w=2)
                    w_f = ly.w_f_kwn
                ops.conv(a_in, w_f, bias, None, act=ly.act, out_nchw=c.heads[ly.name], **kw)
                continue
            out = c.acts[i + 1]
            if ly.norm == L.NORM_NONE:
                ops.conv(a_in, w_f, bias, out, act=ly.act, out_reflect=ly.out_halo > 0, **kw)
                continue
            ops.conv(a_in, w_f, ly.conv.bias if ly.fwd_bias else None, c.yraw[i], **kw)
            nm = ly.norm_mod
            res = c.acts[ly.residual] if ly.residual is not None else None
            if ly.norm == L.NORM_COND_INSTANCE:
                gam, bet = c.cin[i]
                sc, sh = nm.scale_conv[0], nm.shift_conv[0]
                ops.cin_affine_fwd(c.z, sc.weight, sc.bias, sh.weight, sh.bias, gam, bet)
                ops.norm_fwd(c.yraw[i], out, c.nst[i], mode=ly.norm, act=ly.act, gamma=gam, beta=bet, residual=res,
                             eps=nm.eps)
            elif ly.norm == L.NORM_INSTANCE:
                ops.norm_fwd(c.yraw[i], out, c.nst[i], mode=ly.norm, act=ly.act, gamma=nm.scale, beta=nm.shift,
                             residual=res, eps=nm.eps)
            else:  # batch norm (training mode; running stats live in one [2, c] buffer)
                if sync_bn is None:
                    ops.norm_fwd(c.yraw[i], out, c.nst[i], mode=ly.norm, act=ly.act, gamma=nm.weight, beta=nm.bias,
                                 bn_running=running_pair(nm), eps=nm.eps, momentum=nm.momentum)
                else:
                    ops.norm_fwd(c.yraw[i], out, c.nst[i], mode=ly.norm, act=ly.act, gamma=nm.weight, beta=nm.bias,
                                 bn_running=running_pair(nm), eps=nm.eps, momentum=nm.momentum, phase=1)
                    sync_bn(c.nst[i].ws[:2 * out.c])
                    ops.norm_fwd(c.yraw[i], out, c.nst[i], mode=ly.norm, act=ly.act, gamma=nm.weight, beta=nm.bias,
                                 bn_running=running_pair(nm), eps=nm.eps, momentum=nm.momentum, phase=2,
                                 world_size=sync_bn.world_size)
                nm.num_batches_tracked += 1
        return c.heads

    # ---- backward ----------------------------------------------------------------------------
    def backward(self, c, seeds, want_dx=False, want_dw=True, want_dz=False, sync_bn=None, top_grad=None):
        """seeds: {head layer name: True} -- the seed gradient planes c.dyraw[idx] of those heads have been
        filled by the caller (loss kernels / pack_nchw).  Parameter gradients accumulate into the arena.
        top_grad: gradient plane w.r.t. the LAST activation (plans without a head: the standalone residual blocks).
        Returns the input-gradient plane (ring = input halo) if want_dx; want_dx="pair" returns both contributions of
        the input (dgrad of the first layer, residual branch) for the caller to sum."""
        A = self.arena
        pending = {}      # act index -> (dy plane, dy2 plane)
        if top_grad is not None:
            pending[len(self.layers)] = (top_grad, None)
        if want_dz and c.dz is not None:
            c.dz.zero_()
        for i in range(len(self.layers) - 1, -1, -1):
            ly = self.layers[i]
            a_in = c.acts[ly.src]
            dyr = c.dyraw[i]
            if ly.head:
                if ly.name not in seeds:
                    continue
                if want_dw and ly.use_bias:
                    ops.channel_sum(dyr, ly.cout, A.g(ly.conv.bias))
            else:
                if (i + 1) not in pending:
                    continue        # no gradient reaches this layer (e.g. logvar branch)
                dy, dy2 = pending.pop(i + 1)
                nm = ly.norm_mod
                d_res = self._dres(c, ly.residual) if ly.residual is not None else None
                y = c.acts[i + 1] if ly.act != L.ACT_NONE else None
                if ly.norm == L.NORM_NONE:
                    ops.norm_bwd(dy, dyr, c.nst[i], mode=L.NORM_NONE, act=ly.act, y=y, dy2=dy2,
                                 d_beta=A.g(ly.conv.bias) if (want_dw and ly.use_bias) else None, defer_channel=OFFCHAIN_SMALL)
                elif ly.norm == L.NORM_COND_INSTANCE:
                    gam, bet = c.cin[i]
                    ops.norm_bwd(dy, dyr, c.nst[i], mode=ly.norm, act=ly.act, y=y, x=c.yraw[i], gamma=gam, dy2=dy2,
                                 d_res=d_res)
                    if want_dw or want_dz:
                        sc, sh = nm.scale_conv[0], nm.shift_conv[0]
                        if want_dw:
                            tg = (A.g(sc.weight), A.g(sc.bias), A.g(sh.weight), A.g(sh.bias))
                        else:
                            tg = self._scratch_cin(sc)
                        # parameter / noise gradients of the two 1x1 convs: nothing on the chain reads them
                        f = functools.partial(ops.cin_affine_bwd, c.z, sc.weight, sh.weight, gam, bet, c.nst[i].sums, *tg,
                                              c.dz if want_dz else None)
                        ops.off_chain(f) if OFFCHAIN_SMALL else f()
                elif ly.norm == L.NORM_INSTANCE:
                    ops.norm_bwd(dy, dyr, c.nst[i], mode=ly.norm, act=ly.act, y=y, x=c.yraw[i], gamma=nm.scale, dy2=dy2,
                                 d_res=d_res, d_gamma=A.g(nm.scale) if want_dw else None,
                                 d_beta=A.g(nm.shift) if want_dw else None, defer_channel=OFFCHAIN_SMALL)
                else:
                    kw = dict(mode=ly.norm, act=ly.act, y=y, x=c.yraw[i], gamma=nm.weight, dy2=dy2, d_res=d_res,
                              d_gamma=A.g(nm.weight) if want_dw else None, d_beta=A.g(nm.bias) if want_dw else None)
                    if sync_bn is None:
                        ops.norm_bwd(dy, dyr, c.nst[i], **kw)
                    else:
                        ops.norm_bwd(dy, dyr, c.nst[i], phase=1, **kw)
                        sync_bn(c.nst[i].ws[:2 * dyr.c])
                        ops.norm_bwd(dy, dyr, c.nst[i], phase=2, world_size=sync_bn.world_size, **kw)
                if ly.residual is not None:
                    d0, d1 = pending.get(ly.residual, (None, None))
                    pending[ly.residual] = (d_res, d1) if d0 is None else (d0, d_res)
            # weight gradient
            if want_dw:
                ops.off_chain(self._wgrad_fn(ly, a_in, dyr, A.g(ly.conv.weight), c.s2d_dw2[i] if (ly.s2d and c.s2d_cp) else None))
            # data gradient
            if ly.src > 0 or want_dx:
                gin = self._gact(c, ly.src, i)
                ih, iw = c.dims[ly.src]
                if ly.head1:
                    ops.head1_dgrad(dyr, ly.conv.weight, gin, ly.pad)
                elif ly.transposed:
                    ops.conv(dyr, ly.w_d, None, gin, mode=L.CONV_FWD, kh=ly.k, kw=ly.k, stride=ly.stride, pad=ly.pad,
                             cout=ly.cin, out_h=ih, out_w=iw, cin=ly.cout)
                elif (ly.head_kwn_d and a_in.halo == ly.pad and gin.halo == ly.pad and gin.c * gin.t.element_size() == 16
                      and ops.tail_kwn_eligible(dyr.c, ly.k, ly.cin, iw, dyr.dtype)):
                    ops.conv(dyr, ly.w_d_kwn, None, gin, mode=L.CONV_DGRAD, kh=ly.k, kw=ly.k, stride=1, pad=ly.pad,
                             ring=a_in.halo, cout=ly.cin, out_h=ih, out_w=iw, cin=ly.cout, fold_w=2)
                else:
                    ops.conv(dyr, ly.w_d, None, gin, mode=L.CONV_DGRAD, kh=ly.k, kw=ly.k, stride=ly.stride, pad=ly.pad,
                             ring=a_in.halo, cout=ly.cin, out_h=ih, out_w=iw, cin=ly.cout,
                             fold_w=ly.fold_out)
                d0, d1 = pending.get(ly.src, (None, None))
                assert d0 is None or d1 is None, "more than two gradient contributions for one activation"
                pending[ly.src] = (gin, d1) if d0 is None else (gin, d0)
        if want_dw or want_dz:
            ops.off_chain_join()        # the arena (and dz) is complete when backward() returns (in stream order)
        if want_dx == "pair":
            return pending.get(0, (None, None))
        return pending.get(0, (None, None))[0] if want_dx else None

    @staticmethod
    def _wgrad_fn(ly, a_in, dyr, dw, dw2=None):
        """the weight gradient of one layer: off the norm-backward / dgrad chain (ops.off_chain)"""
        if ly.head1:
            return lambda: ops.head1_wgrad(dyr, a_in, dw, ly.pad)
        if dw2 is not None:
            def f():
                ops.conv_wgrad(dyr, a_in, dw2, kh=3, kw=3, stride=1, pad=1, pa=ly.cout, qb=a_in.c)
                ops.s2d_unfold_add(dw2, dw, a_in.s2d)
            return f
        if ly.transposed:
            return lambda: ops.conv_wgrad(a_in, dyr, dw, kh=ly.k, kw=ly.k, stride=ly.stride, pad=ly.pad, pa=ly.cin,
                                          qb=ly.cout)
        return lambda: ops.conv_wgrad(dyr, a_in, dw, kh=ly.k, kw=ly.k, stride=ly.stride, pad=ly.pad, pa=ly.cout,
                                      qb=ly.cin, fold=1 if ly.fold_in else (2 if ly.fold_out else 0))

    def _scratch_cin(self, sc):
        key = ("cin_scratch", sc.weight.shape)
        if key not in self._ctx_cache:
            dev = self.arena.device
            self._ctx_cache[key] = (torch.zeros_like(sc.weight), torch.zeros_like(sc.bias),
                                    torch.zeros_like(sc.weight), torch.zeros_like(sc.bias))
        return self._ctx_cache[key]
