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


def test_kernel_eligibility_predicates():
    """the host-side geometry predicates that decide (per context) which planes / packings the engine prepares must agree
    with the conditions stated in include/dtg_b200.h for dtg_conv fold_w = 2 and the flat-raster DGRAD"""
    from dtg_b200 import ops
    bf, f32 = torch.bfloat16, torch.float32
    # 7x7 tail / head gradient with the filter column in GEMM-N: 32 stored channels, cout <= 4, width divides 128
    assert ops.tail_kwn_eligible(32, 7, 3, 64, bf) and ops.tail_kwn_eligible(32, 7, 1, 128, bf) and ops.tail_kwn_eligible(32, 7, 3, 32, f32)
    assert not ops.tail_kwn_eligible(32, 7, 3, 128, f32)        # two 114 KB patch stages do not fit
    assert not ops.tail_kwn_eligible(32, 7, 3, 96, bf)          # width must divide 128
    assert not ops.tail_kwn_eligible(32, 7, 5, 64, bf)          # (kw, cout) > 28 GEMM columns
    assert not ops.tail_kwn_eligible(48, 7, 3, 64, bf)          # 96-byte pixels: no matching swizzle
    assert not ops.tail_kwn_eligible(32, 4, 3, 64, bf)          # even kernel: not a 'same' convolution
    # flat-raster data gradient of the residual stack: 80 x 32 x 32 x 128 bf16 with ring 1
    assert ops.flat_dgrad_eligible(80, 32, 32, 128, 128, 1, bf)
    assert not ops.flat_dgrad_eligible(80, 32, 32, 128, 128, 0, bf)         # no ring: the ordinary tilings are exact
    assert not ops.flat_dgrad_eligible(80, 64, 64, 128, 128, 1, bf)         # 128 + 2 * 67 rows > one 256-row TMA box
    assert not ops.flat_dgrad_eligible(3, 16, 16, 128, 128, 1, bf)          # 3 * 18 * 18 pixels not divisible by 8
    assert not ops.flat_dgrad_eligible(80, 32, 32, 128, 128, 1, f32)        # eight 25 KB patch units do not fit
    assert not ops.flat_dgrad_eligible(80, 32, 32, 32, 32, 1, bf)           # 64-byte pixels: not full channel chunks


def test_n3_extension_module_trees():
    """N3 on the host: img_size = 64 * 2^k adds k stride-2 stages to LatentEncoder and keeps the reference's
    keys at 64; honor_n_blocks builds range(n_blocks) residual blocks, the default ignores n_blocks like the reference"""
    from dtg_b200 import networks
    bn = networks.get_norm_layer("batch")
    e64 = networks.LatentEncoder(16, 6, 32, norm_layer=bn)
    assert [k for k in e64.state_dict() if k.endswith("weight") and "conv_modules" in k][-2:] == ["conv_modules.11.weight", "conv_modules.12.weight"]
    assert tuple(e64.conv_modules[11].weight.shape) == (256, 256, 4, 4) and e64.n_extra == 0
    e256 = networks.LatentEncoder(16, 6, 32, norm_layer=bn, img_size=256)
    assert e256.n_extra == 2 and tuple(e256.conv_modules[11].weight.shape) == (256, 256, 3, 3)
    assert tuple(e256.conv_modules[17].weight.shape) == (256, 256, 4, 4)
    with pytest.raises(ValueError):
        networks.LatentEncoder(16, 6, 32, norm_layer=bn, img_size=96)
    count = lambda net: len([m for m in net.model if type(m).__name__.endswith("ResnetBlock")])
    assert count(networks.ResnetGenerator(3, 3, 32, n_blocks=9)) == 3                       # networks.py:225
    assert count(networks.CINResnetGenerator(16, 3, 3, 32, n_blocks=9)) == 3                # networks.py:173
    assert count(networks.ResnetGenerator(3, 3, 32, n_blocks=6, honor_n_blocks=True)) == 6
    g0 = networks.CINResnetGenerator(16, 3, 3, 32, n_blocks=0, honor_n_blocks=True)
    assert count(g0) == 0 and type(g0.model[10]).__name__ == "ConvTranspose2d"
