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
import os

import torch
import torch.distributed as dist


class _SyncBN(object):
    """All-reduce (SUM) of the per-channel BatchNorm partial sums between the two phases of dtg_norm_fwd / dtg_norm_bwd.

    The messages are tiny (<= 2 * 256 floats) and sit on the critical path of the encoder / latent-discriminator chains
    (34 exchanges per step), so what matters is latency, not bandwidth.  With symmetric memory (torch.distributed.
    _symmetric_memory: every rank's scratch buffer is peer-mapped over NVLink) the exchange is ONE one-shot kernel --
    barrier on the signal pads, every rank loads all peers' buffers and sums them in rank order (bit-identical on every
    replica) -- instead of an NCCL launch with its protocol handshake; it is CUDA-graph capturable.  If symmetric memory
    cannot be set up (or DTG_SYMM_BN=0) the exchange falls back to ncclAllReduce."""

    def __init__(self, group, world_size):
        self.group, self.world_size = group, world_size
        self._bufs = {}
        self._gname = None
        ok = os.environ.get("DTG_SYMM_BN", "1") != "0" and dist.get_backend(group) == "nccl"
        if ok:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                self._sm = symm_mem
                pg = group if group is not None else dist.group.WORLD
                self._gname = pg.group_name
                torch.ops.symm_mem.one_shot_all_reduce_out      # noqa: B018  (AttributeError -> NCCL path)
            except Exception:
                ok = False
        if dist.get_backend(group) == "nccl":
            # every rank must take the same path: agree on it (a rank without the API would otherwise miss the
            # collective rendezvous of the scratch buffers)
            flag = torch.tensor([1 if ok else 0], device="cuda", dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            ok = bool(int(flag))
        if ok:
            try:
                if hasattr(self._sm, "enable_symm_mem_for_group"):
                    self._sm.enable_symm_mem_for_group(self._gname)
                probe = torch.ones(4, dtype=torch.float32, device="cuda")
                b = self._buf(probe)
                b.copy_(probe)
                torch.ops.symm_mem.one_shot_all_reduce_out(b, "sum", self._gname, probe)
                ok = abs(float(probe[0]) - world_size) < 1e-6
            except Exception:
                ok = False
            flag = torch.tensor([1 if ok else 0], device="cuda", dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            ok = bool(int(flag))
        self.symm = ok

    def _buf(self, sums):
        """peer-mapped scratch of this call site (keyed by the caller's buffer); created collectively, in issue order,
        during the eager warm-up steps that precede graph capture"""
        key = (sums.data_ptr(), sums.numel())
        b = self._bufs.get(key)
        if b is None:
            b = self._sm.empty(sums.numel(), dtype=torch.float32, device=sums.device)
            self._sm.rendezvous(b, self._gname)
            self._bufs[key] = b
        return b

    def __call__(self, sums):
        """sums: fp32 [2*c] device tensor of per-channel partial sums; reduced in place"""
        if self.symm:
            b = self._buf(sums)
            b.copy_(sums)
            torch.ops.symm_mem.one_shot_all_reduce_out(b, "sum", self._gname, sums)
            return
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
