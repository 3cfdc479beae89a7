"""AugmentedCycleGAN with the reference's constructor, attributes, method names and return dicts
(/root/reference/augmented_cyclegan/model.py:337-794), whose ``train_instance`` is ONE fused,
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

from . import _lib as L
from . import engine, networks, ops

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
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
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

    def step(self, grad_scale=1.0):
        ops.step_increment(self.step_dev)
        for name, net in self.nets:
            ex = net._ex
            a = ex.arena
            n = a.active_count
            sq = self.scalars[S_SQ[name]:S_SQ[name] + 1]
            ops.grad_sumsq(a.grad[:n], grad_scale, sq, self.red_ws)
            ops.adam_clip(a.flat[:n], a.grad[:n], a.m[:n], a.v[:n], self.hyper, sq, self.step_dev, grad_scale)
            ex.repack()

    def state_dict(self):
        return {"param_groups": self.param_groups, "step": int(self.step_dev.item()),
                "state": {name: {"exp_avg": net._ex.arena.m.clone(), "exp_avg_sq": net._ex.arena.v.clone()}
                          for name, net in self.nets}}

    def load_state_dict(self, sd):
        self.param_groups[0].update(sd["param_groups"][0])
        self.step_dev.fill_(sd["step"])
        for name, net in self.nets:
            net._exec().arena.m.copy_(sd["state"][name]["exp_avg"])
            net._ex.arena.v.copy_(sd["state"][name]["exp_avg_sq"])
        self.sync_hyper()


class AugmentedCycleGAN(object):
    """Augmented cycle gan (drop-in for model.py:337)."""

    def __init__(self, opt, testing=False):
        self.old_lr = opt.lr
        opt.use_sigmoid = opt.no_lsgan
        self.opt = opt
        if opt.no_lsgan or opt.stoch_enc or opt.use_dropout:
            raise NotImplementedError("dtg_b200: no_lsgan / stoch_enc / use_dropout are not implemented "
                                      "(defaults of options.py:65-71 are)")
        if not torch.cuda.is_available():
            raise RuntimeError("dtg_b200: AugmentedCycleGAN needs a CUDA (sm_100a) device; there is no CPU fallback")
        gpu = [0]    # networks always live on the current CUDA device; one process per GPU
        self.netG_A_B = networks.define_stochastic_G(nlatent=opt.nlatent, input_nc=opt.input_nc,
                                                     output_nc=opt.output_nc, ngf=opt.ngf,
                                                     which_model_netG=opt.which_model_netG, norm=opt.norm,
                                                     use_dropout=opt.use_dropout, gpu_ids=gpu)
        self.netG_B_A = networks.define_G(input_nc=opt.output_nc, output_nc=opt.input_nc, ngf=opt.ngf,
                                          which_model_netG=opt.which_model_netG, norm=opt.norm,
                                          use_dropout=opt.use_dropout, gpu_ids=gpu)
        enc_input_nc = opt.output_nc
        if opt.enc_A_B:
            enc_input_nc += opt.input_nc
        self.netE_B = networks.define_E(nlatent=opt.nlatent, input_nc=enc_input_nc, nef=opt.nef, norm='batch',
                                        gpu_ids=gpu)
        self.netE_B._inactive = ("enc_logvar.weight", "enc_logvar.bias")   # no gradient
        self.netD_A = networks.define_D_A(input_nc=opt.input_nc, ndf=32, which_model_netD=opt.which_model_netD,
                                          norm=opt.norm, use_sigmoid=opt.use_sigmoid, gpu_ids=gpu)
        self.netD_B = networks.define_D_B(input_nc=opt.output_nc, ndf=opt.ndf, which_model_netD=opt.which_model_netD,
                                          norm=opt.norm, use_sigmoid=opt.use_sigmoid, gpu_ids=gpu)
        self.netD_z_B = networks.define_LAT_D(nlatent=opt.nlatent, ndf=opt.ndf, use_sigmoid=opt.use_sigmoid,
                                              gpu_ids=gpu)
        dev = next(self.netG_A_B.parameters()).device
        self.device = dev
        self.scalars = torch.zeros(N_SCALARS, dtype=torch.float32, device=dev)
        self.scalars_host = torch.zeros(N_SCALARS, dtype=torch.float32).pin_memory()
        self.red_ws = torch.zeros(1024, dtype=torch.float32, device=dev)
        mk = lambda nets, lr: FusedAdam(nets, lr, (opt.beta1, 0.999), opt.max_gnorm, self.scalars, self.red_ws)
        # optimizer grouping and learning rates: model.py:379-389
        self.optimizer_G_A = mk([("netG_B_A", self.netG_B_A)], opt.lr)
        self.optimizer_G_B = mk([("netG_A_B", self.netG_A_B), ("netE_B", self.netE_B)], opt.lr)
        self.optimizer_D_A = mk([("netD_A", self.netD_A)], opt.lr / 5.)
        self.optimizer_D_B = mk([("netD_B", self.netD_B), ("netD_z_B", self.netD_z_B)], opt.lr / 5.)
        self.criterionGAN = functools.partial(criterion_GAN, use_sigmoid=opt.use_sigmoid)
        self.criterionCycle = torch.nn.functional.l1_loss
        self.dp = None                 # parallel.DataParallelPlan when running one process per GPU
        self._graph = None
        self._static = None
        if not testing:
            with open("%s/nets.txt" % opt.expr_dir, 'w') as nets_f:
                for net in (self.netG_A_B, self.netG_B_A, self.netD_A, self.netD_B, self.netD_z_B, self.netE_B):
                    networks.print_network(net, nets_f)

    # ---- helpers -------------------------------------------------------------------------------
    def _nets(self):
        return OrderedDict([("netG_A_B", self.netG_A_B), ("netG_B_A", self.netG_B_A), ("netE_B", self.netE_B),
                            ("netD_A", self.netD_A), ("netD_B", self.netD_B), ("netD_z_B", self.netD_z_B)])

    def prepare(self):
        for net in self._nets().values():
            net._exec()
        for o in (self.optimizer_G_A, self.optimizer_G_B, self.optimizer_D_A, self.optimizer_D_B):
            o.sync_hyper()

    @staticmethod
    def _head_idx(ex, name):
        for i, ly in enumerate(ex.layers):
            if ly.name == name:
                return i
        raise KeyError(name)

    # ---- the fused step ------------------------------------------------------------------------
    def _step_device(self, real_A, real_B, prior_z_B):
        """Everything of train_instance that runs on the device (capturable)."""
        o = self.opt
        n, _, h, w = real_A.shape
        nz = o.nlatent
        GAB, GBA, E = self.netG_A_B._ex, self.netG_B_A._ex, self.netE_B._ex
        DA, DB, DZ = self.netD_A._ex, self.netD_B._ex, self.netD_z_B._ex
        sc, ws = self.scalars, self.red_ws
        dp = self.dp
        sync_bn = dp.sync_bn if dp is not None else None
        gs = 1.0 / dp.world_size if dp is not None else 1.0     # all-reduce SUM -> mean of shard gradients
        z_prior = prior_z_B.reshape(n, nz)
        i_mu = self._head_idx(E, "mu")

        # F1 fake_B = G_A_B(real_A, prior_z)                                   model.py:404
        c1 = GAB.new_ctx(n, h, w, "f1")
        ops.pack_nchw(real_A, c1.acts[0], 0)
        c1.z.copy_(z_prior)
        fake_B = GAB.forward(c1)["out"]
        # F2 fake_A = G_B_A(real_B)                                            model.py:407
        c2 = GBA.new_ctx(n, h, w, "f2")
        ops.pack_nchw(real_B, c2.acts[0], 0)
        fake_A = GBA.forward(c2)["out"]
        # F3 mu_z_realB = E_B(cat(fake_A, real_B))                             model.py:409-411
        c3 = E.new_ctx(n, h, w, "f3")
        if o.enc_A_B:
            ops.pack_nchw(fake_A, c3.acts[0], 0)
            ops.pack_nchw(real_B, c3.acts[0], o.input_nc)
        else:
            ops.pack_nchw(real_B, c3.acts[0], 0)
        mu_realB = E.forward(c3, sync_bn)["mu"]                # [n, nz, 1, 1]; post_z_realB (stoch_enc=False)
        ops.loss_l1(mu_realB, mu_realB, 0.0, False, sc, -1, S_KLD, None, ws)      # KLD_z_B, mu_min, mu_max

        # ---- D pass (model.py:423-452): fake.detach() and real as one 2N batch for the IN discriminators
        cdA = DA.new_ctx(2 * n, h, w, "d")
        ops.pack_nchw(fake_A, cdA.acts[0].batch_slice(0, n), 0)
        ops.pack_nchw(real_A, cdA.acts[0].batch_slice(n, 2 * n), 0)
        pA = DA.forward(cdA)["out"]
        iA = self._head_idx(DA, "out")
        ops.loss_lsgan(pA[:n], 0.0, 0.5, sc, S_DFA, S_PFA_D, cdA.dyraw[iA].batch_slice(0, n), ws)
        ops.loss_lsgan(pA[n:], 1.0, 0.5, sc, S_DTA, S_PTA, cdA.dyraw[iA].batch_slice(n, 2 * n), ws)
        cdB = DB.new_ctx(2 * n, h, w, "d")
        ops.pack_nchw(fake_B, cdB.acts[0].batch_slice(0, n), 0)
        ops.pack_nchw(real_B, cdB.acts[0].batch_slice(n, 2 * n), 0)
        pB = DB.forward(cdB)["out"]
        iB = self._head_idx(DB, "out")
        ops.loss_lsgan(pB[:n], 0.0, 0.5, sc, S_DFB, S_PFB_D, cdB.dyraw[iB].batch_slice(0, n), ws)
        ops.loss_lsgan(pB[n:], 1.0, 0.5, sc, S_DTB, S_PTB, cdB.dyraw[iB].batch_slice(n, 2 * n), ws)
        iZ = self._head_idx(DZ, "out")
        cz1 = DZ.new_ctx(n, 1, 1, "d1")                       # batch-norm net: separate calls, reference order
        ops.pack_nchw(mu_realB, cz1.acts[0], 0)
        p1 = DZ.forward(cz1, sync_bn)["out"]
        ops.loss_lsgan(p1, 0.0, 0.5, sc, S_DPZ, -1, cz1.dyraw[iZ], ws)
        cz2 = DZ.new_ctx(n, 1, 1, "d2")
        ops.pack_nchw(prior_z_B.reshape(n, nz, 1, 1), cz2.acts[0], 0)
        p2 = DZ.forward(cz2, sync_bn)["out"]
        ops.loss_lsgan(p2, 1.0, 0.5, sc, S_DQZ, -1, cz2.dyraw[iZ], ws)

        self.optimizer_D_A.zero_grad()
        self.optimizer_D_B.zero_grad()
        ar = dp.allreduce_arena if dp is not None else (lambda arena: None)   # async, overlaps later backward
        DA.backward(cdA, {"out": True})
        ar(DA.arena)
        DB.backward(cdB, {"out": True})
        ar(DB.arena)
        if o.z_gan:
            DZ.backward(cz1, {"out": True}, sync_bn=sync_bn)
            DZ.backward(cz2, {"out": True}, sync_bn=sync_bn)
        ar(DZ.arena)
        if dp is not None:
            dp.wait()
        self.optimizer_D_A.step(gs)
        self.optimizer_D_B.step(gs)

        # ---- G / E pass with the UPDATED discriminators (model.py:457-515)
        cgA = DA.new_ctx(n, h, w, "g")
        ops.pack_nchw(fake_A, cgA.acts[0], 0)
        ops.loss_lsgan(DA.forward(cgA)["out"], 1.0, 1.0, sc, S_GA, S_PFA, cgA.dyraw[iA], ws)
        cgB = DB.new_ctx(n, h, w, "g")
        ops.pack_nchw(fake_B, cgB.acts[0], 0)
        ops.loss_lsgan(DB.forward(cgB)["out"], 1.0, 1.0, sc, S_GB, S_PFB, cgB.dyraw[iB], ws)
        cgZ = DZ.new_ctx(n, 1, 1, "g")
        ops.pack_nchw(mu_realB, cgZ.acts[0], 0)
        ops.loss_lsgan(DZ.forward(cgZ, sync_bn)["out"], 1.0, 1.0 if o.z_gan else 0.0, sc, S_GZ, -1, cgZ.dyraw[iZ], ws)
        # F13 rec_A = G_B_A(fake_B)
        c13 = GBA.new_ctx(n, h, w, "f13")
        ops.pack_nchw(fake_B, c13.acts[0], 0)
        rec_A = GBA.forward(c13)["out"]
        iGo = self._head_idx(GBA, "out")
        ops.loss_l1(rec_A, real_A, o.lambda_A, True, sc, S_CYCA, -1, c13.dyraw[iGo], ws)
        # F14 mu_z_fakeB = E_B(cat(real_A, fake_B))
        c14 = E.new_ctx(n, h, w, "f14")
        if o.enc_A_B:
            ops.pack_nchw(real_A, c14.acts[0], 0)
            ops.pack_nchw(fake_B, c14.acts[0], o.input_nc)
        else:
            ops.pack_nchw(fake_B, c14.acts[0], 0)
        mu_fakeB = E.forward(c14, sync_bn)["mu"]
        ops.loss_l1(mu_fakeB, prior_z_B.reshape(n, nz, 1, 1), o.lambda_z_B, False, sc, S_CYCZ, -1, c14.dyraw[i_mu], ws)
        # F15 rec_B = G_A_B(fake_A, post_z_realB)
        c15 = GAB.new_ctx(n, h, w, "f15")
        ops.pack_nchw(fake_A, c15.acts[0], 0)
        c15.z.copy_(mu_realB.reshape(n, nz))
        rec_B = GAB.forward(c15)["out"]
        iGAo = self._head_idx(GAB, "out")
        ops.loss_l1(rec_B, real_B, o.lambda_B, True, sc, S_CYCB, -1, c15.dyraw[iGAo], ws)

        self.optimizer_G_A.zero_grad()
        self.optimizer_G_B.zero_grad()
        g15 = GAB.backward(c15, {"out": True}, want_dx=True, want_dz=True)          # d fake_A (halo 3), dz
        g14 = E.backward(c14, {"mu": True}, want_dx=True, sync_bn=sync_bn)           # channels 3..5: d fake_B
        g13 = GBA.backward(c13, {"out": True}, want_dx=True)                         # d fake_B (halo 3)
        g12 = DZ.backward(cgZ, {"out": True}, want_dx=True, want_dw=False, sync_bn=sync_bn)   # d post_z
        g11 = DB.backward(cgB, {"out": True}, want_dx=True, want_dw=False)           # d fake_B
        g10 = DA.backward(cgA, {"out": True}, want_dx=True, want_dw=False)           # d fake_A
        # d mu_z_realB = D_z dgrad + dz of F15's CIN projections -> seed of F3's mu head
        ops.grad_gather([g12], [0], nz, out=c3.dyraw[i_mu], add_nchw=c15.dz)
        g3 = E.backward(c3, {"mu": True}, want_dx=True, sync_bn=sync_bn)             # channels 0..2: d fake_A
        ar(E.arena)
        cB = o.input_nc if o.enc_A_B else 0
        ops.grad_gather([g14, g13, g11], [cB, 0, 0], o.output_nc, out=c1.dyraw[iGAo], tanh_y=fake_B)
        GAB.backward(c1, {"out": True})
        ar(GAB.arena)
        ops.grad_gather([g15, g10, g3], [0, 0, 0], o.input_nc, out=c2.dyraw[iGo], tanh_y=fake_A)
        GBA.backward(c2, {"out": True})
        ar(GBA.arena)
        if dp is not None:
            dp.wait()
        self.optimizer_G_A.step(gs)
        self.optimizer_G_B.step(gs)
        return OrderedDict([('real_A', real_A), ('fake_B', fake_B), ('rec_A', rec_A),
                            ('real_B', real_B), ('fake_A', fake_A), ('rec_B', rec_B)])

    def _check_inputs(self, real_A, real_B, prior_z_B):
        for t in (real_A, real_B, prior_z_B):
            if not (t.is_cuda and t.dtype == torch.float32):
                raise ValueError("dtg_b200: train_instance expects float32 CUDA tensors (like the reference after .cuda())")
        if real_A.shape[2] != 64 or real_A.shape[3] != 64:
            raise ValueError("dtg_b200: AugmentedCycleGAN.train_instance needs 64x64 inputs, exactly like the reference "
                             "(LatentEncoder yields [N, nlatent] only at 64x64)")
        return real_A.contiguous(), real_B.contiguous(), prior_z_B.contiguous()

    def _report(self):
        """one packed device->host copy, then the reference's three OrderedDicts (model.py:518-537)"""
        self.scalars_host.copy_(self.scalars, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        s = self.scalars_host.tolist()
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

    def train_instance(self, real_A, real_B, prior_z_B, use_graph=False, report=True):
        """model.py:402-539.  use_graph=True replays a CUDA graph of the whole step (captured on first use
        for this batch shape; inputs are copied into static buffers).  report=False skips the device->host
        read and returns (None, visuals, None)."""
        real_A, real_B, prior_z_B = self._check_inputs(real_A, real_B, prior_z_B)
        self.prepare()
        if use_graph:
            visuals = self._graph_step(real_A, real_B, prior_z_B)
        else:
            visuals = self._step_device(real_A, real_B, prior_z_B)
        if not report:
            return None, visuals, None
        losses, gnorms = self._report()
        if self.opt.monitor_gnorm:
            return losses, visuals, gnorms
        return losses, visuals

    def _graph_step(self, real_A, real_B, prior_z_B):
        key = tuple(real_A.shape)
        if self._graph is None or self._graph[0] != key:
            st = [torch.empty_like(real_A), torch.empty_like(real_B), torch.empty_like(prior_z_B)]
            for s, t in zip(st, (real_A, real_B, prior_z_B)):
                s.copy_(t)
            # warm-up on a side stream (allocations, lazy attribute setup), restoring all state afterwards
            snap = self._snapshot()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._step_device(*st)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self._restore(snap)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                vis = self._step_device(*st)
            self._restore(snap)
            self._graph = (key, g, st, vis)
        _, g, st, vis = self._graph
        for s, t in zip(st, (real_A, real_B, prior_z_B)):
            s.copy_(t, non_blocking=True)
        g.replay()
        vis = OrderedDict(vis)
        vis['real_A'], vis['real_B'] = real_A, real_B
        return vis

    def _snapshot(self):
        snap = []
        for net in self._nets().values():
            a = net._ex.arena
            bufs = {k: v.clone() for k, v in net.state_dict().items() if "running" in k or "num_batches" in k}
            snap.append((a.flat.clone(), a.m.clone(), a.v.clone(), bufs))
        steps = [o.step_dev.clone() for o in (self.optimizer_G_A, self.optimizer_G_B, self.optimizer_D_A, self.optimizer_D_B)]
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
        for o, s in zip((self.optimizer_G_A, self.optimizer_G_B, self.optimizer_D_A, self.optimizer_D_B), steps):
            o.step_dev.copy_(s)

    # ---- inference helpers (model.py:606-733), forward-only through the same plans ---------------
    def _fwd(self, net, heads, *inputs, z=None):
        with torch.no_grad():
            return net._call(heads, inputs[0], z, *inputs[1:])

    def predict_A(self, real_B):
        return self._fwd(self.netG_B_A, ("out",), real_B)

    def predict_B(self, real_A, z_B):
        return self._fwd(self.netG_A_B, ("out",), real_A, z=z_B)

    def _encode(self, a, b):
        ins = (a, b) if self.opt.enc_A_B else (b,)
        mu, logvar = self._fwd(self.netE_B, ("mu", "logvar"), *ins)
        return mu.reshape(mu.size(0), -1), logvar.reshape(logvar.size(0), -1)

    def predict_enc_params(self, real_A, real_B):
        mu, logvar = self._encode(real_A, real_B)
        return (mu,)

    def generate_cycle(self, real_A, real_B, prior_z_B):
        fake_B = self.predict_B(real_A, prior_z_B)
        fake_A = self.predict_A(real_B)
        rec_A = self.predict_A(fake_B)
        mu, _ = self._encode(fake_A, real_B)
        rec_B = self.predict_B(fake_A, mu.reshape(mu.size(0), mu.size(1), 1, 1))
        return OrderedDict([('real_A', real_A), ('fake_B', fake_B), ('rec_A', rec_A),
                            ('real_B', real_B), ('fake_A', fake_A), ('rec_B', rec_B)])

    def generate_multi(self, real_A, multi_prior_z_B):
        size = real_A.size()
        num = multi_prior_z_B.size(0) // real_A.size(0)
        multi_real_A = real_A.unsqueeze(1).repeat(1, num, 1, 1, 1).view(size[0] * num, size[1], size[2], size[3])
        return self.predict_B(multi_real_A, multi_prior_z_B)

    def generate_cycle_B_multi(self, real_B, multi_prior_z_B):
        fake_A = self.predict_A(real_B)
        size = real_B.size()
        num = multi_prior_z_B.size(0) // real_B.size(0)
        multi_fake_A = fake_A.unsqueeze(1).repeat(1, num, 1, 1, 1).view(size[0] * num, size[1], size[2], size[3])
        return fake_A, self.predict_B(multi_fake_A, multi_prior_z_B)

    def inference_multi(self, real_A, real_B):
        size = real_A.size()
        num = real_B.size(0)
        multi_real_A = real_A.unsqueeze(1).repeat(1, num, 1, 1, 1).view(size[0] * num, size[1], size[2], size[3])
        fake_A = self.predict_A(real_B)
        mu, _ = self._encode(fake_A, real_B)
        post_z_B = mu.reshape(mu.size(0), mu.size(1), 1, 1)
        return self.predict_B(multi_real_A, post_z_B.repeat(size[0], 1, 1, 1))

    # ---- bookkeeping (model.py:735-794) ---------------------------------------------------------
    def update_learning_rate(self):
        lrd = self.opt.lr / self.opt.niter_decay
        lr = self.old_lr - lrd
        for o in (self.optimizer_D_A, self.optimizer_G_A, self.optimizer_D_B, self.optimizer_G_B):
            for param_group in o.param_groups:
                param_group['lr'] = lr
            o.sync_hyper()
        print('update learning rate: %f -> %f' % (self.old_lr, lr))
        self.old_lr = lr

    def save(self, chk_name):
        chk_path = os.path.join(self.opt.expr_dir, chk_name)
        checkpoint = {k: net.state_dict() for k, net in self._nets().items()}
        for k in ("optimizer_D_A", "optimizer_G_A", "optimizer_D_B", "optimizer_G_B"):
            checkpoint[k] = getattr(self, k).state_dict()
        torch.save(checkpoint, chk_path)

    def load(self, chk_path):
        checkpoint = torch.load(chk_path)
        for k, net in self._nets().items():
            net.load_state_dict(checkpoint[k])
            net._exec().repack()
        for k in ("optimizer_D_A", "optimizer_G_A", "optimizer_D_B", "optimizer_G_B"):
            if isinstance(checkpoint[k].get("state", None), dict) and "step" in checkpoint[k]:
                getattr(self, k).load_state_dict(checkpoint[k])

    def eval(self):
        for net in self._nets().values():
            net.eval()

    def train(self):
        for net in self._nets().values():
            net.train()
