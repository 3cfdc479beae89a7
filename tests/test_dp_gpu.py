"""Numerical data-parallel parity on two real GPUs: 2 ranks x 4 samples with synchronised BatchNorm and
the NCCL gradient all-reduce inside the fused step must reproduce 1 rank x 8 samples.  Runs tools/dp_parity.py under
torchrun in a subprocess; skipped on boxes with fewer than two GPUs (run it with `gpurun --gpus 2`).

Tolerance.  The instance-norm networks (D_A, D_B) shard exactly: their gradients agree to 1e-5 rel-L2 (measured 5e-8 ..
9e-7; SURVEY 9.2 measured <= 1e-6 for fp32 on CPU).  The BatchNorm-coupled networks (E_B, D_z_B and, through post_z /
mu_z, both generators) are ill-conditioned at this initialisation: the SINGLE-GPU step itself moves by 2e-3 .. 4e-3 (tf32,
batch 8) when the samples are merely fed in reversed order (tools/perm_noise.py; only fp32 summation order differs), so
for those the bound is 3x that permutation noise floor, measured in the same run.  Per-replica BatchNorm statistics give
0.12 .. 2.8 (0.09 .. 1.9), which the second test requires to be detected."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

from tolerances import record

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(extra):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "dp_parity.py")] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    return json.loads(line)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("graph", [False, True])
def test_two_ranks_match_one_rank(graph):
    res = _run(["--precision", "tf32", "--batch", "8"] + (["--graph"] if graph else []))
    record("dp_parity", **res)
    assert res["replicas_identical"]
    for n, e in res["grad_rel"].items():
        bound = 1e-5 if n in ("netD_A", "netD_B") else max(1e-5, 3.0 * res["perm_noise_grad_rel"][n])
        assert e < bound, ("grad", n, e, bound)
    for n in ("netD_A", "netD_B"):
        assert res["weight_rel"][n] < 1e-6, ("weights", n, res["weight_rel"][n])
    for k, e in res["loss_abs"].items():
        assert e < 1e-4, ("loss", k, e)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_unsynchronised_batchnorm_is_detected():
    """the same comparison with per-replica BatchNorm statistics must FAIL the bound (rel-L2 0.09 .. 1.9)"""
    res = _run(["--precision", "tf32", "--batch", "8", "--no-sync-bn"])
    record("dp_parity_nosync", **res)
    for n in ("netG_B_A", "netE_B", "netD_z_B"):
        assert res["grad_rel"][n] > 10.0 * max(1e-5, 3.0 * res["perm_noise_grad_rel"][n]), (n, res["grad_rel"][n])
