"""CPU-side checks (no GPU): the C-ABI library loads and exports every entry point declared in include/dtg_b200.h, the
ctypes structs mirror the header, and the one-process-per-GPU data-parallel plan (sharding, gradient-arena all-reduce,
synchronised batch-norm sums) behaves on a world_size-2 gloo group."""
import ctypes as C
import os
import re
import socket
import sys
import types

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import dtg  # noqa: E402,F401
from dtg_b200 import _lib  # noqa: E402


def _header():
    return open(os.path.join(ROOT, "include", "dtg_b200.h")).read()


def test_library_exports_every_declared_symbol():
    hdr = _header()
    declared = sorted(set(re.findall(r"^\s*(?:int|size_t|unsigned long long)\s+(dtg_\w+)\s*\(", hdr, flags=re.M)))
    assert len(declared) >= 20
    lib = C.CDLL(_lib.LIB_PATH)        # loading needs no GPU; no compute entry point is called here
    for name in declared:
        assert hasattr(lib, name), "libdtg_b200.so does not export %s" % name
    assert sorted(_lib.exported_symbols()) == declared, "ctypes signature table and header disagree"
    lib.dtg_version.restype = C.c_int
    assert lib.dtg_version() == int(re.search(r"#define DTG_VERSION (\d+)", hdr).group(1))


def test_ctypes_structs_match_header():
    hdr = _header()

    def fields(struct):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), hdr, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            decl = re.sub(r"^(const\s+)?[\w ]+?[\s\*]+(?=\w+(\s*,|\s*$))", "", decl)
            names += [n.strip(" *") for n in decl.split(",")]
        return names

    for cname, pyt in (("dtg_plane", _lib.Plane), ("dtg_conv_args", _lib.ConvArgs), ("dtg_wgrad_args", _lib.WgradArgs),
                       ("dtg_norm_args", _lib.NormArgs), ("dtg_pack_item", _lib.PackItem), ("dtg_loss_seg", _lib.LossSeg)):
        assert fields(cname) == [f[0] for f in pyt._fields_], cname


def test_no_cpu_fallback():
    from dtg_b200 import model as dmodel
    from oracle import step as ostep
    import argparse
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    opt = argparse.Namespace(**vars(ostep.default_opt()), expr_dir="/tmp", niter_decay=25)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dmodel.AugmentedCycleGAN(opt, testing=True)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dtg_b200 import parallel
    plan = parallel.DataParallelPlan(sync_bn=True)
    ok = plan.world_size == world and plan.rank == rank
    # sharding: rank r owns rows [r*k, (r+1)*k)
    batch = torch.arange(8.0).reshape(8, 1)
    ok &= torch.equal(plan.shard(batch), batch[rank * 4:(rank + 1) * 4])
    # gradient arenas: SUM over ranks of the ACTIVE prefix only, asynchronously, then wait()
    arena = types.SimpleNamespace(grad=torch.full((10,), float(rank + 1)), active_count=6)
    plan.allreduce_arena(arena)
    plan.wait()
    ok &= torch.equal(arena.grad[:6], torch.full((6,), 3.0)) and torch.equal(arena.grad[6:], torch.full((4,), float(rank + 1)))
    # per-network handles: each optimizer step waits for ITS all-reduce only (the lanes of the fused step)
    a1 = types.SimpleNamespace(grad=torch.full((4,), float(rank + 1)), active_count=4)
    a2 = types.SimpleNamespace(grad=torch.full((4,), 10.0 * (rank + 1)), active_count=4)
    h1, h2 = plan.allreduce_arena(a1), plan.allreduce_arena(a2)
    ok &= h1 is not None and h2 is not None and len(plan._pending) == 2
    plan.wait([h2])
    ok &= torch.equal(a2.grad, torch.full((4,), 30.0)) and len(plan._pending) == 1
    plan.wait([h1, None])
    ok &= torch.equal(a1.grad, torch.full((4,), 3.0)) and len(plan._pending) == 0
    # mean of shard gradients = SUM * (1 / world_size), the grad_scale handed to the fused clip+Adam kernel
    ok &= abs(float(arena.grad[0]) / world - 1.5) < 1e-6
    # synchronised batch-norm: per-channel (sum, sumsq) all-reduced in place
    sums = torch.tensor([1.0, 2.0, 3.0, 4.0]) * (rank + 1)
    plan.sync_bn(sums)
    ok &= torch.equal(sums, torch.tensor([3.0, 6.0, 9.0, 12.0])) and plan.sync_bn.world_size == world
    plan2 = parallel.DataParallelPlan(sync_bn=False)
    ok &= plan2.sync_bn is None
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_plan_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
