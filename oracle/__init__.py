"""oracle/ -- TEST INFRASTRUCTURE ONLY (never imported by the product package).

A plain-PyTorch fp32 restatement of the reference's Augmented CycleGAN training
hot path (adrianalbert/domain-transfer-GAN, ``augmented_cyclegan/{modules,
networks,model}.py``).  The reference delegates every arithmetic op to PyTorch
(``nn.Conv2d``, ``F.mse_loss`` ...; unpinned, idioms of torch ~0.3), so the
restatement calls the same ``torch.nn.functional`` primitives on CPU.

Pinning: the reference ships NO tests / golden vectors for this path
(SURVEY.md section 8c).  The oracle is instead pinned against the reference
itself, imported live from ``/root/reference`` in the build container:
``tests/test_oracle.py`` / ``tests/test_trainer_cpu.py`` (the live-reference cases run only
where /root/reference exists) and the committed fixtures under ``tests/golden/`` produced
by ``tests/golden/make_golden.py`` from the live reference.

Modules: ``nets`` / ``functional`` / ``step`` (the six networks; ``train_instance``,
``supervised_train_instance`` and ``StochCycleGAN.train_instance``), ``live_reference``
(imports the unmodified reference), ``loop`` (restatement of the train.py hot loop --
itself UNPINNED, train.py is Python 2 -- and live execution of the reference's batch
iterators and evaluate.py through text shims).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.
"""
from . import functional, nets, step  # noqa: F401
