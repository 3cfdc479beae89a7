"""dtg_b200 -- B200-native (sm_100a) drop-in for the Augmented CycleGAN training hot path of
adrianalbert/domain-transfer-GAN.  Import through the repo-root shim ``import dtg`` (the directory
name contains '-') or via importlib; the package name is ``dtg_b200``.

Host code is Python/PyTorch (device memory, streams, torch.distributed); all arithmetic on the hot
path runs in the hand-written CUDA kernels of ``csrc/`` behind the C ABI of ``include/dtg_b200.h``.
There is no CPU / eager fallback: every op raises if the extension is missing.
"""
from . import _lib, engine, model, modules, networks, ops, parallel, trainer  # noqa: F401
from .engine import set_precision  # noqa: F401
from .model import AugmentedCycleGAN, StochCycleGAN  # noqa: F401

__all__ = ["_lib", "ops", "engine", "modules", "networks", "model", "parallel", "trainer", "set_precision", "AugmentedCycleGAN",
           "StochCycleGAN"]
