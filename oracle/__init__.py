"""oracle/ -- TEST INFRASTRUCTURE ONLY (never imported by the product package).

A plain-PyTorch fp32 restatement of the reference's Augmented CycleGAN training
hot path (adrianalbert/domain-transfer-GAN, ``augmented_cyclegan/{modules,
networks,model}.py``).  The reference delegates every arithmetic op to PyTorch
(``nn.Conv2d``, ``F.mse_loss`` ...; unpinned, idioms of torch ~0.3), so the
restatement calls the same ``torch.nn.functional`` primitives on CPU.

Pinning: the reference ships NO tests / golden vectors for this path
(SURVEY.md section 8c).  The oracle is instead pinned against the reference
itself, imported live from ``/root/reference`` in the build container:
``tests/test_oracle_vs_reference.py`` (runs only where /root/reference exists)
and the committed fixtures under ``tests/golden/`` produced by
``tests/golden/make_golden.py`` from the live reference.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.
"""
from . import functional, nets, step  # noqa: F401
