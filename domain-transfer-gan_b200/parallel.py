"""Data parallelism: one process per GPU (torchrun), full replicas, batch sharded on dim 0.

Replaces the reference's per-forward ``nn.parallel.data_parallel`` (networks.py:193-197, 248-252, ...:
scatter / re-broadcast all parameters / gather on every call, ~53 MB of parameter traffic per step per
extra GPU) with two exchanges per step over NCCL / NVLink: the D-side and G-side gradient
arenas are all-reduced (SUM) as soon as each network's backward has finished, asynchronously with the
remaining backward kernels, and the fused clip+Adam kernel divides by the world size.

BatchNorm in E_B / D_z_B couples samples: with ``sync_bn=True`` (default) the per-channel
(sum, sum-of-squares) and the matching backward sums are all-reduced so that G GPUs x N/G samples match
one GPU x N samples; ``sync_bn=False`` reproduces the per-replica statistics of the reference's own
data_parallel.
"""
import torch.distributed as dist


class _SyncBN(object):
    def __init__(self, group, world_size):
        self.group, self.world_size = group, world_size

    def __call__(self, sums):
        """sums: fp32 [2*c] device tensor of per-channel partial sums; reduced in place"""
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.group)


class DataParallelPlan(object):
    def __init__(self, group=None, sync_bn=True, overlap=True):
        if not dist.is_initialized():
            raise RuntimeError("DataParallelPlan needs torch.distributed to be initialised (torchrun)")
        self.group = group
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.sync_bn = _SyncBN(group, self.world_size) if (sync_bn and self.world_size > 1) else None
        self.overlap = overlap
        self._pending = []

    def shard(self, t):
        """this rank's slice of a global batch (dim 0)"""
        n = t.shape[0]
        assert n % self.world_size == 0, "global batch must divide the world size"
        k = n // self.world_size
        return t[self.rank * k:(self.rank + 1) * k]

    def allreduce_arena(self, arena):
        """start the all-reduce of one network's gradient arena (call right after its last backward)"""
        if self.world_size == 1:
            return None
        g = arena.grad[:arena.active_count]
        if self.overlap:
            h = dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._pending.append(h)
            return h
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
        return None

    def wait(self, handles=None):
        """make the current stream wait for the given all-reduces (default: all outstanding ones)"""
        for w in (list(self._pending) if handles is None else handles):
            if w is not None:
                w.wait()
                if w in self._pending:
                    self._pending.remove(w)

    def allreduce_grads(self, arenas):
        for a in arenas:
            self.allreduce_arena(a)
        self.wait()

    def broadcast_model(self, model, src=0):
        """make every replica start from rank `src`'s weights / buffers / optimizer state"""
        for net in model._nets().values():
            a = net._exec().arena
            for t in (a.flat, a.m, a.v):
                dist.broadcast(t, src=src, group=self.group)
            for k, b in net.state_dict().items():
                if "running" in k:
                    dist.broadcast(b, src=src, group=self.group)
            net._ex.repack()
