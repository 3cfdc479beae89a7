"""AugmentedCycleGAN and StochCycleGAN with the reference's constructors, attributes, method names and
return dicts (/root/reference/augmented_cyclegan/model.py:337-794 and :75-325), whose ``train_instance``
(and ``supervised_train_instance``, model.py:541-604) is ONE fused,
CUDA-graph-capturable pass over the engine plans: 15 network forwards, hand-scheduled backward with
gradient fan-in kernels, fused LSGAN/L1 loss reductions, clip_grad_norm + Adam on flat arenas, and a
single packed device->host read of the 32 reporting scalars (the reference performs >= 23 syncs).

Differences from the reference, all invisible to its callers:
  * D-pass twin forwards (fake, real) of the instance-norm discriminators run as one batch of 2N
    (instance statistics are per sample, so this is exact);
  * the G-pass does not compute the discriminators' (never consumed) weight gradients;
  * biases in front of a mean-removing norm receive an exact-zero gradient instead of fp32 noise.
Unsupported reference flags raise: stoch_enc=True, no_lsgan=True, use_dropout=True.
"""
import functools
import os
from collections import OrderedDict

import torch

from . import networks, ops

# slots of the packed reporting vector (device fp32[32])
S_DFA, S_DTA, S_DFB, S_DTB, S_DPZ, S_DQZ = 0, 1, 2, 3, 4, 5
S_PFA_D, S_PTA, S_PFB_D, S_PTB = 6, 7, 8, 9
S_GA, S_PFA, S_GB, S_PFB, S_GZ = 10, 11, 12, 13, 14
S_CYCA, S_CYCZ, S_CYCB, S_KLD, S_MUMIN, S_MUMAX = 15, 16, 17, 18, 19, 20
S_SQ = {"netG_A_B": 21, "netG_B_A": 22, "netE_B": 23, "netD_A": 24, "netD_B": 25, "netD_z_B": 26}
N_SCALARS = 32


def criterion_GAN(pred, target_is_real, use_sigmoid=True):
    """model.py:56-72 (LSGAN branch) for callers that use it on tensors directly."""
    if use_sigmoid:
        raise NotImplementedError("dtg_b200: LSGAN only")
    t = torch.ones_like(pred) if target_is_real else torch.zeros_like(pred)
    return torch.nn.functional.mse_loss(pred, t)


class FusedAdam(object):
    """clip_grad_norm + torch.optim.Adam over the flat arenas of the networks it owns
    (model.py:379-389, 447-452, 510-515).  Hyper-parameters and the step counter live on the device so a
    captured CUDA graph sees learning-rate changes."""

    def __init__(self, nets, lr, betas, max_gnorm, scalars, red_ws):
        self.nets = nets            # list of (name, network module)
        self.param_groups = [{"lr": lr, "betas": betas, "eps": 1e-8}]
        dev = scalars.device
        self.hyper = torch.tensor([lr, betas[0], betas[1], 1e-8, max_gnorm], dtype=torch.float32, device=dev)
        # one counter per network: torch.optim.Adam counts steps per parameter and skips parameters without a
        # gradient, so a supervised step (model.py:559-562) advances netD_z_B but not netD_B
        self.step_dev = torch.zeros(len(nets), dtype=torch.int32, device=dev)
        self._lr_on_device = lr
        self.scalars, self.red_ws = scalars, red_ws

    def sync_hyper(self):
        lr = self.param_groups[0]["lr"]
        if lr != self._lr_on_device:
            self.hyper[0:1].copy_(torch.tensor([lr], dtype=torch.float32))
            self._lr_on_device = lr

    def zero_grad(self):
        for _, net in self.nets:
            net._exec().arena.grad.zero_()

    def step(self, grad_scale=1.0, only=None):
        """clip + Adam on every network of the group (or those named in `only`: the ones that received a
        gradient in this backward pass)."""
        for i, (name, net) in enumerate(self.nets):
            if only is not None and name not in only:
                continue
            ex = net._ex
            a = ex.arena
            n = a.active_count
            cnt = self.step_dev[i:i + 1]
            ops.step_increment(cnt)
            sq = self.scalars[S_SQ[name]:S_SQ[name] + 1]
            ops.grad_sumsq(a.grad[:n], grad_scale, sq, self.red_ws[ops.lane()])
            ops.adam_clip(a.flat[:n], a.grad[:n], a.m[:n], a.v[:n], self.hyper, sq, cnt, grad_scale)
            ex.repack()

    def _param_slices(self):
        """(net index, arena, offset, numel, shape, active) per parameter in torch.optim order, i.e.
        itertools.chain(net.parameters() ...) of model.py:109-114, 379-389"""
        out = []
        for i, (name, net) in enumerate(self.nets):
            a = net._exec().arena
            byid = {id(p): n for n, p in a.entries}
            for p in net.parameters():
                n = byid[id(p)]
                off = (a.views[n].data_ptr() - a.flat.data_ptr()) // 4
                out.append((i, a, off, p.numel(), p.shape, n not in a.inactive_names))
        return out

    def state_dict(self):
        """torch.optim.Adam.state_dict() layout (per-parameter step / exp_avg / exp_avg_sq in param_groups
        order), so a checkpoint written by save() loads into the reference's optimizers and vice versa
        (model.py:293-313, 750-778)."""
        steps = self.step_dev.tolist()
        state = {}
        sl = self._param_slices()
        for idx, (i, a, off, nel, shape, active) in enumerate(sl):
            if active and steps[i] > 0:
                state[idx] = {"step": torch.tensor(float(steps[i])),
                              "exp_avg": a.m[off:off + nel].view(shape).clone(),
                              "exp_avg_sq": a.v[off:off + nel].view(shape).clone()}
        g = dict(self.param_groups[0])
        g.update(weight_decay=0, amsgrad=False, params=list(range(len(sl))))
        return {"state": state, "param_groups": [g]}

    def load_state_dict(self, sd):
        g = sd["param_groups"][0]
        sl = self._param_slices()
        if len(g["params"]) != len(sl):
            raise ValueError("loaded state dict contains a parameter group that doesn't match the size of "
                             "optimizer's group")
        self.param_groups[0].update({k: g[k] for k in ("lr", "betas", "eps") if k in g})
        b = self.param_groups[0]["betas"]
        self.hyper[1:4].copy_(torch.tensor([b[0], b[1], self.param_groups[0]["eps"]], dtype=torch.float32))
        steps = [0] * len(self.nets)
        for _, net in self.nets:
            net._ex.arena.m.zero_(); net._ex.arena.v.zero_()
        for idx, (i, a, off, nel, shape, active) in zip(g["params"], sl):
            st = sd["state"].get(idx)
            if st is None or not active:
                continue
            steps[i] = max(steps[i], int(st["step"]))
            a.m[off:off + nel].copy_(st["exp_avg"].reshape(-1))
            a.v[off:off + nel].copy_(st["exp_avg_sq"].reshape(-1))
        self.step_dev.copy_(torch.tensor(steps, dtype=torch.int32))
        self.sync_hyper()


class PendingReport(object):
    """losses / gnorms of one step whose device->host copy is still in flight (train_instance(report="defer"))"""

    def __init__(self, model, host, event):
        self.model, self.host, self.event = model, host, event

    def get(self):
        self.event.synchronize()
        return self.model._dicts(self.host.tolist())


class _FusedCycleModel(object):
    """What AugmentedCycleGAN and StochCycleGAN share: device buffers, the eager / CUDA-graph step drivers,
    forward-only helpers and bookkeeping.  Subclasses define NET_NAMES, OPT_NAMES, _step_device, _report."""
    NET_NAMES = ()
    OPT_NAMES = ()

    def _init_common(self, opt):
        self.old_lr = opt.lr
        opt.use_sigmoid = opt.no_lsgan
        self.opt = opt
        if opt.no_lsgan or getattr(opt, "stoch_enc", False) or opt.use_dropout:
            raise NotImplementedError("dtg_b200: no_lsgan / stoch_enc / use_dropout are not implemented "
                                      "(defaults of options.py:65-71 are)")
        if not torch.cuda.is_available():
            raise RuntimeError("dtg_b200: %s needs a CUDA (sm_100a) device; there is no CPU fallback"
                               % type(self).__name__)
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.scalars = torch.zeros(N_SCALARS, dtype=torch.float32, device=self.device)
        self.scalars_host = torch.zeros(N_SCALARS, dtype=torch.float32).pin_memory()
        self.lanes = ops.Lanes(5, self.device)
        self.red_ws = torch.zeros(5, 4096, dtype=torch.float32, device=self.device)    # reduction scratch, per lane (4 x 4 KB)
        self.criterionGAN = functools.partial(criterion_GAN, use_sigmoid=opt.use_sigmoid)
        self.criterionCycle = torch.nn.functional.l1_loss
        self.dp = None                 # parallel.DataParallelPlan when running one process per GPU
        self._graphs = {}

    def _mk_adam(self, names, lr):
        o = self.opt
        return FusedAdam([(k, getattr(self, k)) for k in names], lr, (o.beta1, 0.999), o.max_gnorm,
                         self.scalars, self.red_ws)

    def _write_nets_txt(self, order):
        with open("%s/nets.txt" % self.opt.expr_dir, 'w') as nets_f:
            for k in order:
                networks.print_network(getattr(self, k), nets_f)

    # ---- helpers -------------------------------------------------------------------------------
    def _nets(self):
        return OrderedDict([(k, getattr(self, k)) for k in self.NET_NAMES])

    def _optimizers(self):
        return [getattr(self, k) for k in self.OPT_NAMES]

    def prepare(self):
        for net in self._nets().values():
            net._exec()
        for o in self._optimizers():
            o.sync_hyper()

    @staticmethod
    def _head_idx(ex, name):
        for i, ly in enumerate(ex.layers):
            if ly.name == name:
                return i
        raise KeyError(name)

    def _dp_env(self):
        dp = self.dp
        sync_bn = dp.sync_bn if dp is not None else None
        gs = 1.0 / dp.world_size if dp is not None else 1.0     # all-reduce SUM -> mean of shard gradients
        ar = dp.allreduce_arena if dp is not None else (lambda arena: None)   # async, overlaps later backward
        return dp, sync_bn, gs, ar

    def _rws(self):
        return self.red_ws[ops.lane()]

    def _read_scalars(self):
        """one packed device->host copy of the reporting vector"""
        self.scalars_host.copy_(self.scalars, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.scalars_host.tolist()

    def _report(self):
        return self._dicts(self._read_scalars())

    def _report_deferred(self):
        """enqueue the packed device->host copy behind the step and return a handle; PendingReport.get() waits for
        THAT copy only, so the caller can issue the next step first and read this one's losses while it runs"""
        k = self._defer_k = (getattr(self, "_defer_k", -1) + 1) % 4
        if not hasattr(self, "_defer_host"):
            self._defer_host = [torch.zeros(N_SCALARS, dtype=torch.float32).pin_memory() for _ in range(4)]
        host = self._defer_host[k]
        host.copy_(self.scalars, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        return PendingReport(self, host, ev)

    def _check_common(self, real_A, real_B, prior_z_B, free_prior=False):
        """free_prior: prior_z_B may have its own batch size (the supervised step only feeds it to D_z_B, and the
        reference's trainer passes the unsupervised batch's prior next to a shorter supervised batch, train.py:211-216)"""
        for t in (real_A, real_B, prior_z_B):
            if not (t.is_cuda and t.dtype == torch.float32):
                raise ValueError("dtg_b200: train_instance expects float32 CUDA tensors (like the reference after .cuda())")
        o = self.opt
        n = real_A.shape[0]
        if (real_A.dim() != 4 or real_B.dim() != 4 or real_A.shape[1] != o.input_nc or real_B.shape[1] != o.output_nc
                or real_B.shape[0] != n or real_A.shape[2:] != real_B.shape[2:]
                or (prior_z_B.shape[0] != n and not free_prior)
                or prior_z_B.numel() != prior_z_B.shape[0] * o.nlatent):
            raise ValueError("dtg_b200: expected real_A [N,%d,H,W], real_B [N,%d,H,W], prior_z_B [N,%d,1,1]"
                             % (o.input_nc, o.output_nc, o.nlatent))
        return real_A.contiguous(), real_B.contiguous(), prior_z_B.contiguous()

    def _run(self, kind, fn, inputs, use_graph):
        self.prepare()
        if use_graph:
            return self._graph_step(kind, fn, inputs)
        return fn(*inputs)

    def train_instance(self, real_A, real_B, prior_z_B, use_graph=False, report=True):
        """model.py:402-539 / :126-208.  use_graph=True replays a CUDA graph of the whole step (captured on
        first use for this batch shape; inputs are copied into static buffers).  report=False skips the
        device->host read and returns (None, visuals, None); report="defer" returns (PendingReport, visuals, None):
        the read is enqueued behind the step and resolved by PendingReport.get() -> (losses, gnorms), at most four
        steps later.
        The visuals (fake_B, rec_A, fake_A, rec_B) are VIEWS of the plan's static head buffers: they are valid until the
        next forward of the same network (the reference returns fresh tensors); clone what must outlive the step."""
        ins = self._check_inputs(real_A, real_B, prior_z_B)
        visuals = self._run("train", self._step_device, ins, use_graph)
        visuals = OrderedDict(visuals)
        visuals['real_A'], visuals['real_B'] = ins[0], ins[1]
        if not report:
            return None, visuals, None
        if report == "defer":
            return self._report_deferred(), visuals, None
        losses, gnorms = self._report()
        if self.opt.monitor_gnorm:
            return losses, visuals, gnorms
        return losses, visuals

    def _graph_step(self, kind, fn, inputs):
        key = (kind,) + tuple(tuple(t.shape) for t in inputs)
        if key not in self._graphs:
            st = [t.clone() for t in inputs]
            # warm-up on a side stream (allocations, lazy attribute setup), restoring all state afterwards
            snap = self._snapshot()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    fn(*st)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self._restore(snap)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = fn(*st)
            self._restore(snap)
            self._graphs[key] = (g, st, out)
        g, st, out = self._graphs[key]
        for s, t in zip(st, inputs):
            s.copy_(t, non_blocking=True)
        g.replay()
        return out

    def _snapshot(self):
        snap = []
        for net in self._nets().values():
            a = net._ex.arena
            bufs = {k: v.clone() for k, v in net.state_dict().items() if "running" in k or "num_batches" in k}
            snap.append((a.flat.clone(), a.m.clone(), a.v.clone(), bufs))
        steps = [o.step_dev.clone() for o in self._optimizers()]
        return snap, steps

    def _restore(self, snapshot):
        snap, steps = snapshot
        for net, (flat, m, v, bufs) in zip(self._nets().values(), snap):
            a = net._ex.arena
            a.flat.copy_(flat); a.m.copy_(m); a.v.copy_(v)
            sd = net.state_dict()
            for k, t in bufs.items():
                sd[k].copy_(t)
            net._ex.repack()
        for o, s in zip(self._optimizers(), steps):
            o.step_dev.copy_(s)

    # ---- forward-only helpers shared by both models ------------------------------------------------
    def _fwd(self, net, heads, *inputs, z=None):
        """forward-only unless an input asks for a gradient: evaluate.py:70-123 (variational_ubo) differentiates
        predict_B(real_A, z_B) with respect to z_B, so that call stays on the autograd path of networks._NetFn"""
        need = torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in inputs + (z,))
        with torch.set_grad_enabled(need):
            return net._call(heads, inputs[0], z, *inputs[1:])

    def _z(self, z_B):
        return z_B

    def predict_A(self, real_B):
        return self._fwd(self.netG_B_A, ("out",), real_B)

    def predict_B(self, real_A, z_B):
        return self._fwd(self.netG_A_B, ("out",), real_A, z=self._z(z_B))

    def _prior(self, like):
        """real_B.data.new(N, nlatent, 1, 1).normal_(0, 1) (model.py:227,249,669)"""
        return torch.randn(like.size(0), self.opt.nlatent, 1, 1, device=like.device, dtype=torch.float32)

    @staticmethod
    def _repeat(x, num):
        size = x.size()
        return x.unsqueeze(1).repeat(1, num, 1, 1, 1).view(size[0] * num, size[1], size[2], size[3])

    def generate_multi(self, real_A, multi_prior_z_B):
        num = multi_prior_z_B.size(0) // real_A.size(0)
        return self.predict_B(self._repeat(real_A, num), multi_prior_z_B)

    def generate_cycle_B_multi(self, real_B, multi_prior_z_B):
        fake_A = self.predict_A(real_B)
        num = multi_prior_z_B.size(0) // real_B.size(0)
        return fake_A, self._fwd(self.netG_A_B, ("out",), self._repeat(fake_A, num), z=multi_prior_z_B)

    def _noisy(self, fake_A, std):
        """model.py:253-257 / :628-631"""
        noise_std = std / 127.5
        return torch.clamp(fake_A + torch.randn_like(fake_A) * noise_std, -1, 1)

    # ---- bookkeeping (model.py:282-325, 735-794) ----------------------------------------------------
    def update_learning_rate(self):
        lrd = self.opt.lr / self.opt.niter_decay
        lr = self.old_lr - lrd
        for o in self._optimizers():
            for param_group in o.param_groups:
                param_group['lr'] = lr
            o.sync_hyper()
        print('update learning rate: %f -> %f' % (self.old_lr, lr))
        self.old_lr = lr

    def save(self, chk_name):
        chk_path = os.path.join(self.opt.expr_dir, chk_name)
        checkpoint = {k: net.state_dict() for k, net in self._nets().items()}
        for k in self.OPT_NAMES:
            checkpoint[k] = getattr(self, k).state_dict()
        torch.save(checkpoint, chk_path)

    def load(self, chk_path):
        """Loads a checkpoint written by save() or by the reference's save() (model.py:293-313, 750-778; the
        state_dicts have identical keys).  The reference's torch.optim.Adam state is per parameter in param_groups
        order: FusedAdam.load_state_dict scatters it into the flat m / v arenas."""
        checkpoint = torch.load(chk_path, map_location=self.device)
        for k, net in self._nets().items():
            net.load_state_dict(checkpoint[k])
            net._exec().repack()
        for k in self.OPT_NAMES:
            getattr(self, k).load_state_dict(checkpoint[k])

    def eval(self):
        """model.py:298-303 flips the nn.Module flags and so does this.  The fused plans do not have a running-statistics
        inference mode: the BatchNorm networks (E_B, D_z_B) keep normalising with batch statistics (and keep updating
        running_mean / running_var) after eval().  No code path of the reference ever calls eval() (train.py,
        evaluate.py and test.py all run the networks in training mode), so the drop-in behaviour is identical there; a
        caller that relies on eval() semantics gets a warning instead of silently different numbers."""
        import warnings
        if any(isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)) for net in self._nets().values()
               for m in net.modules()):
            warnings.warn("dtg_b200: eval() does not switch the fused BatchNorm layers to running statistics; E_B / D_z_B "
                          "keep using batch statistics, exactly as in the reference's own (training-mode) evaluation paths")
        for net in self._nets().values():
            net.eval()

    def train(self):
        for net in self._nets().values():
            net.train()


class AugmentedCycleGAN(_FusedCycleModel):
    """Augmented cycle gan (drop-in for model.py:337)."""
    NET_NAMES = ("netG_A_B", "netG_B_A", "netE_B", "netD_A", "netD_B", "netD_z_B")
    OPT_NAMES = ("optimizer_D_A", "optimizer_G_A", "optimizer_D_B", "optimizer_G_B")

    def __init__(self, opt, testing=False):
        self._init_common(opt)
        gpu = [0]    # networks always live on the current CUDA device; one process per GPU
        # N3 extensions, both off unless the option is present: opt.n_blocks (int) is honoured as the number
        # of res-blocks (the reference ignores it, networks.py:173/225 -> 3); opt.encoder_grid_size = 64 * 2^k gives
        # E_B k extra stride-2 stages so that the model runs above 64x64 (the reference cannot)
        nb = getattr(opt, "n_blocks", None)
        self.grid = int(getattr(opt, "encoder_grid_size", 64) or 64)
        self.netG_A_B = networks.define_stochastic_G(nlatent=opt.nlatent, input_nc=opt.input_nc,
                                                     output_nc=opt.output_nc, ngf=opt.ngf,
                                                     which_model_netG=opt.which_model_netG, norm=opt.norm,
                                                     use_dropout=opt.use_dropout, gpu_ids=gpu, n_blocks=nb)
        self.netG_B_A = networks.define_G(input_nc=opt.output_nc, output_nc=opt.input_nc, ngf=opt.ngf,
                                          which_model_netG=opt.which_model_netG, norm=opt.norm,
                                          use_dropout=opt.use_dropout, gpu_ids=gpu, n_blocks=nb)
        enc_input_nc = opt.output_nc
        if opt.enc_A_B:
            enc_input_nc += opt.input_nc
        self.netE_B = networks.define_E(nlatent=opt.nlatent, input_nc=enc_input_nc, nef=opt.nef, norm='batch',
                                        gpu_ids=gpu, img_size=self.grid)
        self.netE_B._inactive = ("enc_logvar.weight", "enc_logvar.bias")   # no gradient
        self.netD_A = networks.define_D_A(input_nc=opt.input_nc, ndf=32, which_model_netD=opt.which_model_netD,
                                          norm=opt.norm, use_sigmoid=opt.use_sigmoid, gpu_ids=gpu)
        self.netD_B = networks.define_D_B(input_nc=opt.output_nc, ndf=opt.ndf, which_model_netD=opt.which_model_netD,
                                          norm=opt.norm, use_sigmoid=opt.use_sigmoid, gpu_ids=gpu)
        self.netD_z_B = networks.define_LAT_D(nlatent=opt.nlatent, ndf=opt.ndf, use_sigmoid=opt.use_sigmoid,
                                              gpu_ids=gpu)
        # optimizer grouping and learning rates: model.py:379-389
        self.optimizer_G_A = self._mk_adam(("netG_B_A",), opt.lr)
        self.optimizer_G_B = self._mk_adam(("netG_A_B", "netE_B"), opt.lr)
        self.optimizer_D_A = self._mk_adam(("netD_A",), opt.lr / 5.)
        self.optimizer_D_B = self._mk_adam(("netD_B", "netD_z_B"), opt.lr / 5.)
        if not testing:
            self._write_nets_txt(("netG_A_B", "netG_B_A", "netD_A", "netD_B", "netD_z_B", "netE_B"))

    # ---- the fused step ------------------------------------------------------------------------
    def _step_device(self, real_A, real_B, prior_z_B):
        """Everything of train_instance that runs on the device (capturable).

        The step is issued as a DAG over five lanes (ops.Lanes: parallel CUDA streams, parallel branches of the
        captured graph).  Lanes 0 and 1 carry the generator chains: first forward, then -- without waiting for the
        discriminators -- the cycle forward and its backward, finally the first forward's backward.  Lanes 3 and 4
        carry D_B and D_A (D pass, Adam, generator-side adversarial forward + dgrad), lane 2 the launch-bound small
        networks (E_B, D_z_B); their kernels fill the SMs the big convolutions leave idle, and the weight gradients
        of every backward run on the lanes' companion streams.  Each network's forwards / backwards stay on one
        lane or are ordered by events (BatchNorm running statistics, gradient accumulation into the arena); scratch
        buffers are per lane.  The issue order below is a valid serial schedule and equals the reference order of
        model.py:402-515 up to commuting independent operations."""
        o = self.opt
        n, _, h, w = real_A.shape
        nz = o.nlatent
        GAB, GBA, E = self.netG_A_B._ex, self.netG_B_A._ex, self.netE_B._ex
        DA, DB, DZ = self.netD_A._ex, self.netD_B._ex, self.netD_z_B._ex
        sc = self.scalars
        ws = self._rws
        dp, sync_bn, gs, ar = self._dp_env()
        z_prior = prior_z_B.reshape(n, nz)
        z_prior4 = prior_z_B.reshape(n, nz, 1, 1)
        i_mu = self._head_idx(E, "mu")
        iA, iB, iZ = self._head_idx(DA, "out"), self._head_idx(DB, "out"), self._head_idx(DZ, "out")
        iGo, iGAo = self._head_idx(GBA, "out"), self._head_idx(GAB, "out")
        c1, c2, c3 = GAB.new_ctx(n, h, w, "f1"), GBA.new_ctx(n, h, w, "f2"), E.new_ctx(n, h, w, "f3")
        cdA, cdB = DA.new_ctx(2 * n, h, w, "d"), DB.new_ctx(2 * n, h, w, "d")
        cz1, cz2, cgZ = DZ.new_ctx(n, 1, 1, "d1"), DZ.new_ctx(n, 1, 1, "d2"), DZ.new_ctx(n, 1, 1, "g")
        cgA, cgB = DA.new_ctx(n, h, w, "g"), DB.new_ctx(n, h, w, "g")
        c13, c14, c15 = GBA.new_ctx(n, h, w, "f13"), E.new_ctx(n, h, w, "f14"), GAB.new_ctx(n, h, w, "f15")
        r = {}
        wait = (lambda hs: dp.wait(hs)) if dp is not None else (lambda hs: None)
        for opt_ in self._optimizers():      # model.py:443-444, 507-508 (nothing reads .grad in between)
            opt_.zero_grad()
        ln = self.lanes
        ln.begin()

        def f1():       # fake_B = G_A_B(real_A, prior_z)                          model.py:404
            ops.pack_nchw(real_A, c1.acts[0], 0)
            c1.z.copy_(z_prior)
            r["fake_B"] = GAB.forward(c1)["out"]

        def f2():       # fake_A = G_B_A(real_B)                                   model.py:407
            ops.pack_nchw(real_B, c2.acts[0], 0)
            r["fake_A"] = GBA.forward(c2)["out"]

        def f3():       # mu_z_realB = E_B(cat(fake_A, real_B)); KLD_z_B, mu_min, mu_max      model.py:409-421
            if o.enc_A_B:
                ops.pack_nchw(r["fake_A"], c3.acts[0], 0)
                ops.pack_nchw(real_B, c3.acts[0], o.input_nc)
            else:
                ops.pack_nchw(real_B, c3.acts[0], 0)
            r["mu"] = E.forward(c3, sync_bn)["mu"]             # [n, nz, 1, 1]; post_z_realB (stoch_enc=False)
            ops.loss_l1(r["mu"], r["mu"], 0.0, False, sc, -1, S_KLD, None, ws())

        e_f1 = ln.run(0, f1)
        e_f2 = ln.run(1, f2)
        e_f3 = ln.run(2, f3, after=(e_f2,))

        # ---- second generator forwards + their backward (model.py:467-468, 493-494): they need fake_A / fake_B / mu
        # only, NOT the updated discriminators, so they follow the first forwards on lanes 0 / 1 while the whole D pass
        # runs beside them on lanes 2-4
        def g_cyc_0():
            ops.pack_nchw(r["fake_B"], c13.acts[0], 0)          # rec_A = G_B_A(fake_B)
            r["rec_A"] = GBA.forward(c13)["out"]
            ops.loss_l1(r["rec_A"], real_A, o.lambda_A, True, sc, S_CYCA, -1, c13.dyraw[iGo], ws())
            r["g13"] = GBA.backward(c13, {"out": True}, want_dx=True)                        # d fake_B (halo 3)

        def g_cyc_1():
            ops.pack_nchw(r["fake_A"], c15.acts[0], 0)          # rec_B = G_A_B(fake_A, post_z_realB)
            c15.z.copy_(r["mu"].reshape(n, nz))
            r["rec_B"] = GAB.forward(c15)["out"]
            ops.loss_l1(r["rec_B"], real_B, o.lambda_B, True, sc, S_CYCB, -1, c15.dyraw[iGAo], ws())
            r["g15"] = GAB.backward(c15, {"out": True}, want_dx=True, want_dz=True)          # d fake_A (halo 3), dz

        e_b0 = ln.run(0, g_cyc_0)
        e_b1 = ln.run(1, g_cyc_1, after=(e_f3,))

        # ---- D pass (model.py:423-452): fake.detach() and real as one 2N batch for the IN discriminators; then the
        # generator-side adversarial terms with the UPDATED discriminators (model.py:457-464) on the same lane
        def d_pair(ex, c, i, fake, real, s_fake, s_true, s_pf, s_pt):
            ops.pack_nchw(fake, c.acts[0].batch_slice(0, n), 0)
            ops.pack_nchw(real, c.acts[0].batch_slice(n, 2 * n), 0)
            p = ex.forward(c)["out"]
            # 0.5 * (mse(D(fake), 0) + mse(D(real), 1)) and both seed gradients: ONE two-segment reduction (model.py:327-334)
            ops.loss_fused([ops.lsgan_seg(p[:n], 0.0, 0.5, s_fake, s_pf, c.dyraw[i].batch_slice(0, n)),
                            ops.lsgan_seg(p[n:], 1.0, 0.5, s_true, s_pt, c.dyraw[i].batch_slice(n, 2 * n))], sc, ws())
            ex.backward(c, {"out": True})
            r[ex] = ar(ex.arena)

        def step_of(optim, ex, name):
            wait([r.get(ex)])
            optim.step(gs, only=(name,))

        def d_b():
            d_pair(DB, cdB, iB, r["fake_B"], real_B, S_DFB, S_DTB, S_PFB_D, S_PTB)
            step_of(self.optimizer_D_B, DB, "netD_B")
            ops.pack_nchw(r["fake_B"], cgB.acts[0], 0)
            ops.loss_lsgan(DB.forward(cgB)["out"], 1.0, 1.0, sc, S_GB, S_PFB, cgB.dyraw[iB], ws())
            r["g11"] = DB.backward(cgB, {"out": True}, want_dx=True, want_dw=False)          # d fake_B

        def d_a():
            d_pair(DA, cdA, iA, r["fake_A"], real_A, S_DFA, S_DTA, S_PFA_D, S_PTA)
            step_of(self.optimizer_D_A, DA, "netD_A")
            ops.pack_nchw(r["fake_A"], cgA.acts[0], 0)
            ops.loss_lsgan(DA.forward(cgA)["out"], 1.0, 1.0, sc, S_GA, S_PFA, cgA.dyraw[iA], ws())
            r["g10"] = DA.backward(cgA, {"out": True}, want_dx=True, want_dw=False)          # d fake_A

        def d_z():      # batch-norm net: separate calls, reference order
            ops.pack_nchw(r["mu"], cz1.acts[0], 0)
            p_post = DZ.forward(cz1, sync_bn)["out"]
            ops.pack_nchw(z_prior4, cz2.acts[0], 0)
            p_prior = DZ.forward(cz2, sync_bn)["out"]
            ops.loss_fused([ops.lsgan_seg(p_post, 0.0, 0.5, S_DPZ, -1, cz1.dyraw[iZ]),
                            ops.lsgan_seg(p_prior, 1.0, 0.5, S_DQZ, -1, cz2.dyraw[iZ])], sc, ws())
            if o.z_gan:
                DZ.backward(cz1, {"out": True}, sync_bn=sync_bn)
                DZ.backward(cz2, {"out": True}, sync_bn=sync_bn)
            r[DZ] = ar(DZ.arena)
            step_of(self.optimizer_D_B, DZ, "netD_z_B")
            ops.pack_nchw(r["mu"], cgZ.acts[0], 0)
            ops.loss_lsgan(DZ.forward(cgZ, sync_bn)["out"], 1.0, 1.0 if o.z_gan else 0.0, sc, S_GZ, -1, cgZ.dyraw[iZ], ws())

        def e_cyc():    # mu_z_fakeB = E_B(cat(real_A, fake_B)) and its backward (model.py:471-486)
            if o.enc_A_B:
                ops.pack_nchw(real_A, c14.acts[0], 0)
                ops.pack_nchw(r["fake_B"], c14.acts[0], o.input_nc)
            else:
                ops.pack_nchw(r["fake_B"], c14.acts[0], 0)
            mu_fakeB = E.forward(c14, sync_bn)["mu"]
            ops.loss_l1(mu_fakeB, z_prior4, o.lambda_z_B, False, sc, S_CYCZ, -1, c14.dyraw[i_mu], ws())
            r["g14"] = E.backward(c14, {"mu": True}, want_dx=True, sync_bn=sync_bn)          # channels 3..5: d fake_B
            r["g12"] = DZ.backward(cgZ, {"out": True}, want_dx=True, want_dw=False, sync_bn=sync_bn)   # d post_z

        def e_first():
            # d mu_z_realB = D_z dgrad + dz of F15's CIN projections -> seed of F3's mu head
            ops.grad_gather([r["g12"]], [0], nz, out=c3.dyraw[i_mu], add_nchw=c15.dz)
            # with enc_A_B the encoder saw cat(fake_A, real_B): channels 0..2 of its input gradient are d fake_A;
            # without it the encoder saw real_B only and fake_A receives nothing from this path (model.py:409-413)
            r["g3"] = E.backward(c3, {"mu": True}, want_dx=bool(o.enc_A_B), sync_bn=sync_bn)
            r[E] = ar(E.arena)

        e_db = ln.run(3, d_b, after=(e_f1,))
        e_da = ln.run(4, d_a, after=(e_f2,))
        ln.run(2, d_z)
        e_b2a = ln.run(2, e_cyc, after=(e_f1,))
        e_b2b = ln.run(2, e_first, after=(e_b1,))

        def g_last_0():
            cB = o.input_nc if o.enc_A_B else 0
            ops.grad_gather([r["g14"], r["g13"], r["g11"]], [cB, 0, 0], o.output_nc, out=c1.dyraw[iGAo], tanh_y=r["fake_B"])
            GAB.backward(c1, {"out": True})
            r[GAB] = ar(GAB.arena)

        def g_last_1():
            srcs = [r["g15"], r["g10"]] + ([r["g3"]] if o.enc_A_B else [])
            ops.grad_gather(srcs, [0] * len(srcs), o.input_nc, out=c2.dyraw[iGo], tanh_y=r["fake_A"])
            GBA.backward(c2, {"out": True})
            r[GBA] = ar(GBA.arena)

        ln.run(0, g_last_0, after=(e_b2a, e_db, e_b1))     # e_b1: c15's backward wrote G_A_B's gradient arena
        ln.run(1, g_last_1, after=(e_b2b, e_da, e_b0))     # e_b0: c13's backward wrote G_B_A's gradient arena
        ln.run(0, lambda: step_of(self.optimizer_G_B, GAB, "netG_A_B"))
        ln.run(1, lambda: step_of(self.optimizer_G_A, GBA, "netG_B_A"))
        ln.run(2, lambda: step_of(self.optimizer_G_B, E, "netE_B"))
        ln.end()
        return OrderedDict([('real_A', real_A), ('fake_B', r["fake_B"]), ('rec_A', r["rec_A"]),
                            ('real_B', real_B), ('fake_A', r["fake_A"]), ('rec_B', r["rec_B"])])

    def _check_inputs(self, real_A, real_B, prior_z_B, free_prior=False):
        ins = self._check_common(real_A, real_B, prior_z_B, free_prior)
        if real_A.shape[2] != self.grid or real_A.shape[3] != self.grid:
            raise ValueError("dtg_b200: this AugmentedCycleGAN needs %dx%d inputs: the reference's LatentEncoder yields "
                             "[N, nlatent] only at 64x64; build the model with opt.encoder_grid_size = 64 * 2^k for "
                             "larger grids (extension), or use StochCycleGAN, which takes any size"
                             % (self.grid, self.grid))
        return ins

    def _dicts(self, s):
        """the reference's OrderedDicts (model.py:518-537) from the packed reporting vector"""
        losses = OrderedDict([('D_A', 0.5 * (s[S_DFA] + s[S_DTA])), ('G_A', s[S_GA]), ('Cyc_A', s[S_CYCA]),
                              ('Cyc_z_B', s[S_CYCZ]), ('KLD_z_B', s[S_KLD]),
                              ('D_B', 0.5 * (s[S_DFB] + s[S_DTB])), ('G_B', s[S_GB]), ('Cyc_B', s[S_CYCB]),
                              ('D_z_B', 0.5 * (s[S_DPZ] + s[S_DQZ])),
                              ('P_t_A', s[S_PTA]), ('P_f_A', s[S_PFA]), ('P_t_B', s[S_PTB]), ('P_f_B', s[S_PFB])])
        gn = lambda k: s[S_SQ[k]] ** 0.5
        gnorms = OrderedDict([('gnorm_G_A_B', gn("netG_A_B")), ('gnorm_G_B_A', gn("netG_B_A")),
                              ('gnorm_E_B', gn("netE_B")), ('gnorm_D_B', gn("netD_B")),
                              ('gnorm_D_z_B', gn("netD_z_B")), ('gnorm_D_A', gn("netD_A")),
                              ('mu_min', s[S_MUMIN]), ('mu_max', s[S_MUMAX]),
                              ('logvar_min', 0.0), ('logvar_max', 0.0)])
        return losses, gnorms

    # ---- the supervised step (model.py:541-604) -----------------------------------------------------
    def _sup_device(self, real_A, real_B, prior_z_B):
        o = self.opt
        n, _, h, w = real_A.shape
        nz = o.nlatent
        GAB, GBA, E, DZ = self.netG_A_B._ex, self.netG_B_A._ex, self.netE_B._ex, self.netD_z_B._ex
        sc, ws = self.scalars, self.red_ws[0]
        dp, sync_bn, gs, ar = self._dp_env()
        i_mu, iZ = self._head_idx(E, "mu"), self._head_idx(DZ, "out")
        iGAo, iGo = self._head_idx(GAB, "out"), self._head_idx(GBA, "out")
        # mu = E_B(cat(real_A, real_B)); post_z_B = mu                           model.py:543-553
        cE = E.new_ctx(n, h, w, "s")
        if o.enc_A_B:
            ops.pack_nchw(real_A, cE.acts[0], 0)
            ops.pack_nchw(real_B, cE.acts[0], o.input_nc)
        else:
            ops.pack_nchw(real_B, cE.acts[0], 0)
        mu = E.forward(cE, sync_bn)["mu"]
        ops.loss_l1(mu, mu, 0.0, False, sc, -1, S_KLD, None, ws)               # KLD_z_B (:574)
        # latent discriminator update                                            model.py:555-562
        cz1 = DZ.new_ctx(n, 1, 1, "d1")
        ops.pack_nchw(mu, cz1.acts[0], 0)
        ops.loss_lsgan(DZ.forward(cz1, sync_bn)["out"], 0.0, 0.5, sc, S_DPZ, -1, cz1.dyraw[iZ], ws)
        nq = prior_z_B.shape[0]            # its own batch size: discriminate() runs the two forwards separately (:327-334)
        cz2 = DZ.new_ctx(nq, 1, 1, "d2")
        ops.pack_nchw(prior_z_B.reshape(nq, nz, 1, 1), cz2.acts[0], 0)
        ops.loss_lsgan(DZ.forward(cz2, sync_bn)["out"], 1.0, 0.5, sc, S_DQZ, -1, cz2.dyraw[iZ], ws)
        self.optimizer_D_B.zero_grad()
        DZ.backward(cz1, {"out": True}, sync_bn=sync_bn)
        DZ.backward(cz2, {"out": True}, sync_bn=sync_bn)
        ar(DZ.arena)
        if dp is not None:
            dp.wait()
        self.optimizer_D_B.step(gs, only=("netD_z_B",))      # netD_B has no gradient: torch skips it
        # supervised reconstruction + latent GAN term                            model.py:564-581
        cB = GAB.new_ctx(n, h, w, "s")
        ops.pack_nchw(real_A, cB.acts[0], 0)
        cB.z.copy_(mu.reshape(n, nz))
        pred_B = GAB.forward(cB)["out"]
        ops.loss_l1(pred_B, real_B, o.lambda_sup_B, True, sc, S_CYCB, -1, cB.dyraw[iGAo], ws)
        cA = GBA.new_ctx(n, h, w, "s")
        ops.pack_nchw(real_B, cA.acts[0], 0)
        pred_A = GBA.forward(cA)["out"]
        ops.loss_l1(pred_A, real_A, o.lambda_sup_A, True, sc, S_CYCA, -1, cA.dyraw[iGo], ws)
        cgZ = DZ.new_ctx(n, 1, 1, "g")
        ops.pack_nchw(mu, cgZ.acts[0], 0)
        ops.loss_lsgan(DZ.forward(cgZ, sync_bn)["out"], 1.0, 1.0 if o.z_gan else 0.0, sc, S_GZ, -1, cgZ.dyraw[iZ], ws)
        self.optimizer_G_A.zero_grad()
        self.optimizer_G_B.zero_grad()
        GBA.backward(cA, {"out": True})
        ar(GBA.arena)
        GAB.backward(cB, {"out": True}, want_dz=True)
        ar(GAB.arena)
        gz = DZ.backward(cgZ, {"out": True}, want_dx=True, want_dw=False, sync_bn=sync_bn)
        ops.grad_gather([gz], [0], nz, out=cE.dyraw[i_mu], add_nchw=cB.dz)
        E.backward(cE, {"mu": True}, sync_bn=sync_bn)
        ar(E.arena)
        if dp is not None:
            dp.wait()
        self.optimizer_G_A.step(gs)
        self.optimizer_G_B.step(gs)
        return ()

    def supervised_train_instance(self, real_A, real_B, prior_z_B, use_graph=False):
        """model.py:541-604; returns the reference's loss dict (:593-602)."""
        ins = self._check_inputs(real_A, real_B, prior_z_B, free_prior=True)
        self._run("sup", self._sup_device, ins, use_graph)
        s = self._read_scalars()
        gn = lambda k: s[S_SQ[k]] ** 0.5
        return OrderedDict([('S_A', s[S_CYCA]), ('S_B', s[S_CYCB]), ('KLD_z_B', s[S_KLD]),
                            ('D_z_B', 0.5 * (s[S_DPZ] + s[S_DQZ])),
                            ('gnorm_G_A_B', gn("netG_A_B")), ('gnorm_G_B_A', gn("netG_B_A")),
                            ('gnorm_E_B', gn("netE_B")), ('gnorm_D_z_B', gn("netD_z_B"))])

    # ---- inference helpers (model.py:606-733), forward-only through the same plans ---------------
    def _encode(self, a, b):
        ins = (a, b) if self.opt.enc_A_B else (b,)
        mu, logvar = self._fwd(self.netE_B, ("mu", "logvar"), *ins)
        return mu.reshape(mu.size(0), -1), logvar.reshape(logvar.size(0), -1)

    def _post_z(self, a, b):
        mu, _ = self._encode(a, b)
        return mu.reshape(mu.size(0), mu.size(1), 1, 1)

    def predict_enc_params(self, real_A, real_B):
        mu, logvar = self._encode(real_A, real_B)
        return (mu,)

    def generate_cycle(self, real_A, real_B, prior_z_B):
        fake_B = self.predict_B(real_A, prior_z_B)
        fake_A = self.predict_A(real_B)
        rec_A = self.predict_A(fake_B)
        rec_B = self.predict_B(fake_A, self._post_z(fake_A, real_B))
        return OrderedDict([('real_A', real_A), ('fake_B', fake_B), ('rec_A', rec_A),
                            ('real_B', real_B), ('fake_A', fake_A), ('rec_B', rec_B)])

    def generate_noisy_cycle(self, real_B, std):
        """model.py:626-645"""
        fake_A = self.predict_A(real_B)
        noisy_fake_A = self._noisy(fake_A, std)
        return self.predict_B(noisy_fake_A, self._post_z(fake_A, real_B))

    def generate_multi_cycle(self, real_B, steps, from_prior=True):
        """model.py:664-685"""
        images = [real_B]
        B = real_B
        for i in range(steps):
            A = self.predict_A(B)
            z_B = self._prior(real_B) if from_prior else self._post_z(A, B)
            B = self.predict_B(A, z_B)
            images.extend([A, B])
        return images

    def inference_multi(self, real_A, real_B):
        size = real_A.size()
        num = real_B.size(0)
        multi_real_A = self._repeat(real_A, num)
        fake_A = self.predict_A(real_B)
        post_z_B = self._post_z(fake_A, real_B)
        return self.predict_B(multi_real_A, post_z_B.repeat(size[0], 1, 1, 1))


class StochCycleGAN(_FusedCycleModel):
    """Stochastic cycle gan (drop-in for model.py:75): the two generators and the two image discriminators,
    no encoder and no latent discriminator, hence fully convolutional -- the reference-native training step at
    128x128 / 256x256 (BASELINE configs 3 and 4).  ignore_noise=True is the plain CycleGAN of train.py:159-160."""
    NET_NAMES = ("netG_A_B", "netG_B_A", "netD_A", "netD_B")
    OPT_NAMES = ("optimizer_D", "optimizer_G")

    def __init__(self, opt, ignore_noise=False, testing=False):
        self.ignore_noise = ignore_noise
        self._init_common(opt)
        gpu = [0]
        nb = getattr(opt, "n_blocks", None)     # honoured when present (N3 extension); None = the reference's 3 blocks
        self.netG_A_B = networks.define_stochastic_G(nlatent=opt.nlatent, input_nc=opt.input_nc,
                                                     output_nc=opt.output_nc, ngf=opt.ngf,
                                                     which_model_netG=opt.which_model_netG, norm=opt.norm,
                                                     use_dropout=opt.use_dropout, gpu_ids=gpu, n_blocks=nb)
        self.netG_B_A = networks.define_G(input_nc=opt.output_nc, output_nc=opt.input_nc, ngf=opt.ngf,
                                          which_model_netG=opt.which_model_netG, norm=opt.norm,
                                          use_dropout=opt.use_dropout, gpu_ids=gpu, n_blocks=nb)
        self.netD_A = networks.define_D_A(input_nc=opt.input_nc, ndf=32, which_model_netD=opt.which_model_netD,
                                          norm=opt.norm, use_sigmoid=opt.use_sigmoid, gpu_ids=gpu)
        self.netD_B = networks.define_D_B(input_nc=opt.output_nc, ndf=opt.ndf, which_model_netD=opt.which_model_netD,
                                          norm=opt.norm, use_sigmoid=opt.use_sigmoid, gpu_ids=gpu)
        # one Adam for both generators, one for both discriminators: model.py:109-114
        self.optimizer_G = self._mk_adam(("netG_A_B", "netG_B_A"), opt.lr)
        self.optimizer_D = self._mk_adam(("netD_A", "netD_B"), opt.lr / 5.)
        if not testing:
            self._write_nets_txt(("netG_A_B", "netG_B_A", "netD_A", "netD_B"))

    def _z(self, z_B):
        return z_B.mul(0.).add(1.) if self.ignore_noise else z_B          # model.py:128-129, 264-265

    def _check_inputs(self, real_A, real_B, prior_z_B):
        ins = self._check_common(real_A, real_B, prior_z_B)
        h, w = real_A.shape[2:]
        if h % 16 or w % 16 or h < 64 or w < 64:
            raise ValueError("dtg_b200: StochCycleGAN needs H, W >= 64 and divisible by 16 (four stride-2 stages "
                             "and the 4x4 valid conv of Discriminator_edges, networks.py:365-382)")
        return ins

    def _step_device(self, real_A, real_B, prior_z_B):
        """model.py:126-208 on the device (capturable), issued over four lanes like AugmentedCycleGAN._step_device:
        lanes 0 / 1 carry the generator chains, lanes 3 / 4 the D_B / D_A passes."""
        o = self.opt
        n, _, h, w = real_A.shape
        nz = o.nlatent
        GAB, GBA, DA, DB = self.netG_A_B._ex, self.netG_B_A._ex, self.netD_A._ex, self.netD_B._ex
        sc = self.scalars
        ws = self._rws
        dp, sync_bn, gs, ar = self._dp_env()
        iGAo, iGo = self._head_idx(GAB, "out"), self._head_idx(GBA, "out")
        iA, iB = self._head_idx(DA, "out"), self._head_idx(DB, "out")
        z = self._z(prior_z_B).reshape(n, nz)
        c1, c2 = GAB.new_ctx(n, h, w, "f1"), GBA.new_ctx(n, h, w, "f2")
        cdA, cdB = DA.new_ctx(2 * n, h, w, "d"), DB.new_ctx(2 * n, h, w, "d")
        cgA, cgB = DA.new_ctx(n, h, w, "g"), DB.new_ctx(n, h, w, "g")
        c13, c15 = GBA.new_ctx(n, h, w, "f13"), GAB.new_ctx(n, h, w, "f15")
        r = {}
        wait = (lambda hs: dp.wait(hs)) if dp is not None else (lambda hs: None)
        self.optimizer_D.zero_grad()
        self.optimizer_G.zero_grad()
        ln = self.lanes
        ln.begin()

        def f1():       # fake_B = G_A_B(real_A, z)                                 model.py:132
            ops.pack_nchw(real_A, c1.acts[0], 0)
            c1.z.copy_(z)
            r["fake_B"] = GAB.forward(c1)["out"]

        def f2():       # fake_A = G_B_A(real_B)                                    model.py:135
            ops.pack_nchw(real_B, c2.acts[0], 0)
            r["fake_A"] = GBA.forward(c2)["out"]

        e_f1 = ln.run(0, f1)
        e_f2 = ln.run(1, f2)

        # ---- D pass (model.py:137-163): fake.detach() and real as one 2N batch (instance statistics are per sample)
        def d_pair(ex, c, i, fake, real, s_fake, s_true, s_pf, s_pt):
            ops.pack_nchw(fake, c.acts[0].batch_slice(0, n), 0)
            ops.pack_nchw(real, c.acts[0].batch_slice(n, 2 * n), 0)
            p = ex.forward(c)["out"]
            # 0.5 * (mse(D(fake), 0) + mse(D(real), 1)) and both seed gradients: ONE two-segment reduction (model.py:327-334)
            ops.loss_fused([ops.lsgan_seg(p[:n], 0.0, 0.5, s_fake, s_pf, c.dyraw[i].batch_slice(0, n)),
                            ops.lsgan_seg(p[n:], 1.0, 0.5, s_true, s_pt, c.dyraw[i].batch_slice(n, 2 * n))], sc, ws())
            ex.backward(c, {"out": True})
            r[ex] = ar(ex.arena)

        def step_of(optim, ex, name):
            def f():
                wait([r.get(ex)])
                optim.step(gs, only=(name,))
            return f

        def d_b():
            d_pair(DB, cdB, iB, r["fake_B"], real_B, S_DFB, S_DTB, S_PFB_D, S_PTB)
            step_of(self.optimizer_D, DB, "netD_B")()
            ops.pack_nchw(r["fake_B"], cgB.acts[0], 0)
            ops.loss_lsgan(DB.forward(cgB)["out"], 1.0, 1.0, sc, S_GB, S_PFB, cgB.dyraw[iB], ws())
            r["g11"] = DB.backward(cgB, {"out": True}, want_dx=True, want_dw=False)          # d fake_B

        def d_a():
            d_pair(DA, cdA, iA, r["fake_A"], real_A, S_DFA, S_DTA, S_PFA_D, S_PTA)
            step_of(self.optimizer_D, DA, "netD_A")()
            ops.pack_nchw(r["fake_A"], cgA.acts[0], 0)
            ops.loss_lsgan(DA.forward(cgA)["out"], 1.0, 1.0, sc, S_GA, S_PFA, cgA.dyraw[iA], ws())
            r["g10"] = DA.backward(cgA, {"out": True}, want_dx=True, want_dw=False)          # d fake_A

        # ---- cycle forwards + backward (model.py:175-180): independent of the discriminators, so they follow the
        # first forwards on lanes 0 / 1 while the D pass and the adversarial terms run beside them on lanes 3 / 4
        def g_cyc_0():
            ops.pack_nchw(r["fake_B"], c13.acts[0], 0)          # rec_A = G_B_A(fake_B)
            r["rec_A"] = GBA.forward(c13)["out"]
            ops.loss_l1(r["rec_A"], real_A, o.lambda_A, True, sc, S_CYCA, -1, c13.dyraw[iGo], ws())
            r["g13"] = GBA.backward(c13, {"out": True}, want_dx=True)                        # d fake_B

        def g_cyc_1():
            ops.pack_nchw(r["fake_A"], c15.acts[0], 0)          # rec_B = G_A_B(fake_A, z)
            c15.z.copy_(z)
            r["rec_B"] = GAB.forward(c15)["out"]
            ops.loss_l1(r["rec_B"], real_B, o.lambda_B, True, sc, S_CYCB, -1, c15.dyraw[iGAo], ws())
            r["g15"] = GAB.backward(c15, {"out": True}, want_dx=True)                        # d fake_A

        e_b0 = ln.run(0, g_cyc_0)
        e_b1 = ln.run(1, g_cyc_1)
        e_db = ln.run(3, d_b, after=(e_f1,))
        e_da = ln.run(4, d_a, after=(e_f2,))

        def g_last_0():
            ops.grad_gather([r["g13"], r["g11"]], [0, 0], o.output_nc, out=c1.dyraw[iGAo], tanh_y=r["fake_B"])
            GAB.backward(c1, {"out": True})
            r[GAB] = ar(GAB.arena)

        def g_last_1():
            ops.grad_gather([r["g15"], r["g10"]], [0, 0], o.input_nc, out=c2.dyraw[iGo], tanh_y=r["fake_A"])
            GBA.backward(c2, {"out": True})
            r[GBA] = ar(GBA.arena)

        ln.run(0, g_last_0, after=(e_db, e_b1))      # e_b1: c15's backward wrote G_A_B's gradient arena
        ln.run(1, g_last_1, after=(e_da, e_b0))      # e_b0: c13's backward wrote G_B_A's gradient arena
        ln.run(0, step_of(self.optimizer_G, GAB, "netG_A_B"))
        ln.run(1, step_of(self.optimizer_G, GBA, "netG_B_A"))
        ln.end()
        return OrderedDict([('real_A', real_A), ('fake_B', r["fake_B"]), ('rec_A', r["rec_A"]),
                            ('real_B', real_B), ('fake_A', r["fake_A"]), ('rec_B', r["rec_B"])])

    def _dicts(self, s):
        """model.py:193-206"""
        losses = OrderedDict([('D_A', 0.5 * (s[S_DFA] + s[S_DTA])), ('G_A', s[S_GA]), ('Cyc_A', s[S_CYCA]),
                              ('D_B', 0.5 * (s[S_DFB] + s[S_DTB])), ('G_B', s[S_GB]), ('Cyc_B', s[S_CYCB]),
                              ('P_t_A', s[S_PTA]), ('P_f_A', s[S_PFA]), ('P_t_B', s[S_PTB]), ('P_f_B', s[S_PFB])])
        gn = lambda k: s[S_SQ[k]] ** 0.5
        gnorms = OrderedDict([('gnorm_G_A_B', gn("netG_A_B")), ('gnorm_G_B_A', gn("netG_B_A")),
                              ('gnorm_D_B', gn("netD_B")), ('gnorm_D_A', gn("netD_A"))])
        return losses, gnorms

    # ---- forward-only helpers (model.py:210-280) ----------------------------------------------------
    def generate_cycle(self, real_A, real_B, prior_z_B):
        fake_B = self.predict_B(real_A, prior_z_B)
        fake_A = self.predict_A(real_B)
        rec_A = self.predict_A(fake_B)
        rec_B = self.predict_B(fake_A, prior_z_B)
        return OrderedDict([('real_A', real_A), ('fake_B', fake_B), ('rec_A', rec_A),
                            ('real_B', real_B), ('fake_A', fake_A), ('rec_B', rec_B)])

    def generate_multi_cycle(self, real_B, steps):
        images = [real_B]
        B = real_B
        for i in range(steps):
            A = self.predict_A(B)
            B = self.predict_B(A, self._prior(real_B))
            images.extend([A, B])
        return images

    def generate_noisy_cycle(self, real_B, std):
        fake_A = self.predict_A(real_B)
        z_B = self._prior(real_B)
        return self.predict_B(self._noisy(fake_A, std), z_B)

    def generate_multi(self, real_A, multi_prior_z_B):
        return super(StochCycleGAN, self).generate_multi(real_A, multi_prior_z_B)
