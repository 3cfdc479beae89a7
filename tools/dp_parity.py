"""Data-parallel numerical parity on real GPUs: W ranks x (N / W) samples with synchronised BatchNorm
and NCCL gradient all-reduce must reproduce 1 rank x N samples -- every network gradient, the updated weights and the
reported losses.  Launch with torchrun (one rank per GPU); rank 0 prints ONE JSON line.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
        tools/dp_parity.py [--precision tf32|bf16] [--batch 8] [--graph] [--no-sync-bn]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="tf32")
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--graph", action="store_true")
ap.add_argument("--no-sync-bn", action="store_true")
args = ap.parse_args()

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
out = os.fdopen(os.dup(1), "w")      # NCCL prints its banner to fd 1: keep the JSON line on the real stdout
os.dup2(2, 1)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()

import dtg  # noqa: E402,F401
from dtg_b200 import engine, model as dmodel, parallel  # noqa: E402
from oracle import nets as onets, step as ostep  # noqa: E402

engine.set_precision(args.precision)
state = onets.init_model_state(seed=1234, perturb=0.05)
a, b, z = [t.cuda() for t in ostep.synthetic_batch(args.batch, seed=4321)]


def build():
    opt = argparse.Namespace(**vars(ostep.default_opt()), expr_dir="/tmp", niter_decay=25)
    m = dmodel.AugmentedCycleGAN(opt, testing=True)
    for name, net in m._nets().items():
        net.load_state_dict({k: v.clone() for k, v in state[name].items()}, strict=False)
    m.prepare()
    for net in m._nets().values():
        net._ex.repack()
    return m


def snapshot(m):
    g, w = {}, {}
    for name, net in m._nets().items():
        ar = net._ex.arena
        g[name] = ar.grad[:ar.active_count].clone()
        w[name] = ar.flat[:ar.active_count].clone()
    return g, w


# ---- W ranks x N/W samples -------------------------------------------------------------------------------------
mp = build()
mp.dp = parallel.DataParallelPlan(sync_bn=not args.no_sync_bn)
mp.dp.broadcast_model(mp)
k = args.batch // world
sl = slice(rank * k, (rank + 1) * k)
lp, _, gp = mp.train_instance(a[sl].contiguous(), b[sl].contiguous(), z[sl].contiguous(), use_graph=args.graph)
torch.cuda.synchronize()
gpar, wpar = snapshot(mp)
keys = [kk for kk in lp]
lt = torch.tensor([lp[kk] for kk in keys], device="cuda", dtype=torch.float64)
dist.all_reduce(lt)
lt /= world
mp._graphs = {}

# ---- 1 rank x N samples (every rank computes it; rank 0 reports) ------------------------------------------------
m1 = build()
l1, _, g1 = m1.train_instance(a, b, z)
torch.cuda.synchronize()
gone, wone = snapshot(m1)

# ---- noise floor: the same single-GPU step with the samples in reversed order (a mathematically neutral change; only
# floating-point summation order differs).  The BatchNorm-coupled networks are ill-conditioned at this initialisation,
# so this floor -- not fp32 epsilon -- is what a data-parallel run can be expected to reproduce.
m2 = build()
perm = torch.arange(args.batch - 1, -1, -1, device="cuda")
m2.train_instance(a[perm].contiguous(), b[perm].contiguous(), z[perm].contiguous())
torch.cuda.synchronize()
gperm, _ = snapshot(m2)


def rel(x, y):
    return float((x - y).norm() / y.norm().clamp_min(1e-20))


# ---- the BatchNorm exchange in isolation (well conditioned, unlike the encoder at initialisation): one BN2d + ReLU layer,
# forward and backward, W ranks x N/W samples with the (sum, sum-of-squares) / backward-sum exchange between the two kernel
# phases (engine.py forward / backward do exactly this) against one rank x N samples
def bn_unit(sync):
    from dtg_b200 import _lib as L, ops
    g = torch.Generator().manual_seed(99)
    n, c, h = args.batch, 64, 8
    x_all = (torch.randn(n, c, h, h, generator=g) * 1.5 + 0.7).cuda()
    dy_all = torch.randn(n, c, h, h, generator=g).cuda()
    gamma = (torch.rand(c, generator=g) + 0.5).cuda()
    beta = torch.randn(c, generator=g).cuda()

    def run(x, dy, exchange, world_size):
        xp = ops.PlaneT.from_nchw(x, dtype=torch.float32)
        dyp = ops.PlaneT.from_nchw(dy, dtype=torch.float32)
        out = ops.PlaneT(x.shape[0], h, h, c, 0, torch.float32)
        dx = ops.PlaneT(x.shape[0], h, h, c, 0, torch.float32)
        st = ops.NormState(xp)
        run_stats = torch.stack([torch.zeros(c), torch.ones(c)]).cuda()
        dg, db = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
        kw = dict(mode=L.NORM_BATCH, act=L.ACT_RELU, gamma=gamma, beta=beta, bn_running=run_stats)
        bw = dict(mode=L.NORM_BATCH, act=L.ACT_RELU, y=out, x=xp, gamma=gamma, d_gamma=dg, d_beta=db)
        if exchange is None:
            ops.norm_fwd(xp, out, st, **kw)
            ops.norm_bwd(dyp, dx, st, **bw)
        else:
            ops.norm_fwd(xp, out, st, phase=1, **kw)
            exchange(st.ws[:2 * c])
            ops.norm_fwd(xp, out, st, phase=2, world_size=world_size, **kw)
            ops.norm_bwd(dyp, dx, st, phase=1, **bw)
            exchange(st.ws[:2 * c])
            ops.norm_bwd(dyp, dx, st, phase=2, world_size=world_size, **bw)
        torch.cuda.synchronize()
        return out.to_nchw(c), dx.to_nchw(c), dg, db, run_stats

    full = run(x_all, dy_all, None, 1)
    plan = parallel.DataParallelPlan(sync_bn=True)
    ex = plan.sync_bn if sync else (lambda t: None)
    part = run(x_all[sl].contiguous(), dy_all[sl].contiguous(), ex, world if sync else 1)
    dgb = torch.stack([part[2], part[3]])
    dist.all_reduce(dgb)                                     # parameter gradients: summed over ranks like the arenas
    return {"y": rel(part[0], full[0][sl]), "dx": rel(part[1], full[1][sl]), "d_gamma": rel(dgb[0], full[2]),
            "d_beta": rel(dgb[1], full[3]), "running": rel(part[4], full[4])}


res = {"world": world, "precision": args.precision, "batch": args.batch, "graph": bool(args.graph),
       "bn_unit": bn_unit(not args.no_sync_bn),
       "sync_bn": not args.no_sync_bn,
       "grad_rel": {n: rel(gpar[n], gone[n]) for n in gone},
       "perm_noise_grad_rel": {n: rel(gperm[n], gone[n]) for n in gone},
       "weight_rel": {n: rel(wpar[n], wone[n]) for n in wone},
       "loss_abs": {kk: abs(float(lt[i]) - l1[kk]) for i, kk in enumerate(keys)},
       "gnorm_rel": {kk: abs(gp[kk] - g1[kk]) / max(abs(g1[kk]), 1e-12) for kk in g1 if kk.startswith("gnorm")}}
# every replica must hold bit-identical weights after the step
same = True
for n in wpar:
    t = wpar[n].clone()
    dist.broadcast(t, src=0)
    same &= bool(torch.equal(t, wpar[n]))
flag = torch.tensor([1 if same else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
res["replicas_identical"] = bool(int(flag))
if rank == 0:
    out.write(json.dumps(res) + "\n")
    out.flush()
torch.cuda.synchronize()
dist.barrier()
sys.stdout.flush()
sys.stderr.flush()
os._exit(0)
