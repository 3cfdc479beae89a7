"""Numerical data-parallel parity on two real GPUs: 2 ranks x 4 samples with synchronised BatchNorm and
the NCCL gradient all-reduce inside the fused step must reproduce 1 rank x 8 samples.  Runs tools/dp_parity.py under
torchrun in a subprocess; skipped on boxes with fewer than two GPUs (run it with `gpurun --gpus 2`).

What can be held to SURVEY 9.2's 1e-5 and what cannot:
* the instance-norm networks (D_A, D_B) shard exactly: gradients 1e-5 rel-L2 (measured 5e-8 .. 9e-7), updated weights 1e-6;
* every reported loss (forward quantities, including the synchronised-BatchNorm forward of E_B / D_z_B): 1e-4 absolute
  (measured <= 2e-5; per-replica statistics move KLD_z_B / D_z_B / Cyc_z_B by 1e-2 .. 4e-2);
* the BatchNorm exchange itself, in isolation on a well-conditioned layer (BN2d + ReLU forward and backward, `bn_unit`): y, dx,
  d_gamma, d_beta and the running statistics to 1e-5 (per-replica statistics: > 1e-2);
* the gradients of the BatchNorm-COUPLED networks (E_B, D_z_B and, through post_z / mu_z, both generators) at this
  initialisation cannot: they are ill-conditioned in ANY reduced precision -- the reference's own cuDNN-TF32 step differs from
  its fp32 step by 0.24 (G_B_A) / 0.26 (E_B) rel-L2 at batch 8 and by 0.02 .. 0.07 under a mere reversal of the sample order
  (tools/perm_noise_oracle.py, profiles/r2_perm_noise_oracle.txt; ours: the same magnitudes).  For those the test only asks
  that the data-parallel gradient is no further from the single-GPU one than the single-GPU run is from itself under that
  permutation (3x, floor 2e-2), and relies on the three exact checks above to catch a wrong exchange."""
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
    for k, e in res["bn_unit"].items():
        assert e < 1e-5, ("bn_unit", k, e)
    for n in ("netD_A", "netD_B"):
        assert res["grad_rel"][n] < 1e-5, ("grad", n, res["grad_rel"][n])
        assert res["weight_rel"][n] < 1e-6, ("weights", n, res["weight_rel"][n])
    for k, e in res["loss_abs"].items():
        assert e < 1e-4, ("loss", k, e)
    for n in ("netG_A_B", "netG_B_A", "netE_B", "netD_z_B"):
        bound = max(2e-2, 3.0 * res["perm_noise_grad_rel"][n])
        assert res["grad_rel"][n] < bound, ("grad", n, res["grad_rel"][n], bound)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_unsynchronised_batchnorm_is_detected():
    """the same comparison with per-replica BatchNorm statistics must FAIL the exact checks"""
    res = _run(["--precision", "tf32", "--batch", "8", "--no-sync-bn"])
    record("dp_parity_nosync", **res)
    assert res["bn_unit"]["y"] > 1e-2 and res["bn_unit"]["dx"] > 1e-2 and res["bn_unit"]["d_gamma"] > 1e-2
    assert max(res["loss_abs"][k] for k in ("KLD_z_B", "D_z_B", "Cyc_z_B")) > 1e-3
    assert res["grad_rel"]["netD_z_B"] > 0.1                 # BN1d network: well conditioned, 1e-3 when synchronised
    for n in ("netD_A", "netD_B"):                           # no BatchNorm: unaffected
        assert res["grad_rel"][n] < 1e-5
