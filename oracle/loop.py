"""Oracle restatement of the reference's training hot loop (train.py:185-256, 307-313) -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED for the loop itself: train.py is Python 2 (print statements, ``.next()``) and imports cPickle /
skimage, so it cannot be executed in this container; this file restates its control flow line by line, and
tests/test_trainer_cpu.py compares the product loop (dtg_b200.trainer.train_epochs) against it with recording stub
models.  The batch iterators it is driven with ARE pinned: tests exec the reference's own AlignedIterator /
UnalignedIterator source (dataloader.py:60-149, integer-division shim only) and compare index for index.
"""
import torch


def reference_iterators():
    """(AlignedIterator, UnalignedIterator) classes built from the reference's own source text (dataloader.py:60-149);
    the only edits: Python-2 integer division ``/`` -> ``//`` in the two n_batches lines."""
    import numpy as np
    from . import live_reference as lr
    import os
    src = open(os.path.join(lr.REF_DIR, "dataloader.py")).read()
    a = src.index("class AlignedIterator")
    b = src.index("class NumpyDataset")
    txt = src[a:b].replace("self.num_samples / batch_size", "self.num_samples // batch_size")
    txt = txt.replace("self.num_samples / self.batch_size", "self.num_samples // self.batch_size")
    ns = {"np": np, "torch": torch}
    exec(compile(txt, "dataloader.py[60:149]", "exec"), ns)
    return ns["AlignedIterator"], ns["UnalignedIterator"]


def iterate(it):
    """Python-2 style iteration of a reference iterator (``.next()`` until StopIteration)"""
    while True:
        try:
            yield it.next()
        except StopIteration:
            return


def train_loop(model, opt, train_dataset, draw_z, log, sup_train_dataset=None):
    """train.py:185-256, 307-313 with the visualisation / evaluation blocks removed.  `draw_z(n)` stands for
    ``real_A.data.new(n, nlatent, 1, 1).normal_(0, 1)`` (train.py:190), `log(message)` for print_log."""
    total_steps = 0
    for epoch in range(opt.epoch_count, opt.niter + opt.niter_decay + 1):           # :185
        epoch_iter = 0
        for data in iterate(train_dataset):                                          # :189
            real_A, real_B = data['A'], data['B']
            if real_A.size(0) != real_B.size(0):                                     # :191-192
                continue
            prior_z_B = draw_z(real_A.size(0))                                       # :193
            total_steps += opt.batchSize                                             # :195-196
            epoch_iter += opt.batchSize
            out = model.train_instance(real_A, real_B, prior_z_B)                    # :205-208
            losses = out[0]
            gnorms = out[2] if opt.monitor_gnorm else None
            if getattr(opt, "supervised", False):                                    # :211-216
                sup = sup_train_dataset.next()
                sup_losses = model.supervised_train_instance(sup['A'], sup['B'], prior_z_B)
            if total_steps % opt.print_freq == 0:                                    # :243-249
                log(("print", epoch, epoch_iter, dict(losses)))
                if getattr(opt, "supervised", False):
                    log(("print_sup", epoch, epoch_iter, dict(sup_losses)))
                if opt.monitor_gnorm:
                    log(("print_gnorm", epoch, epoch_iter, dict(gnorms)))
        if epoch % opt.save_epoch_freq == 0:                                         # :251-254
            log(("save", epoch, total_steps))
            model.save('latest')
        if epoch > opt.niter:                                                        # :312-313
            model.update_learning_rate()
    return total_steps


def reference_evaluate():
    """The reference's evaluate.py executed live (build container only) with text shims: Python-2 ``print res_str`` ->
    ``print(res_str)``, ``.data[0]`` -> ``.item()`` (0-dim indexing), ``.cuda()`` removed (CPU run), and the model import
    redirected to the live reference model module.  Returns the module namespace (eval_mse_A, variational_ubo, ...)."""
    import os
    import types
    import sys
    from . import live_reference as lr
    _, _, M = lr.load()
    src = open(os.path.join(lr.REF_DIR, "evaluate.py")).read()
    src = src.replace("print res_str", "print(res_str)").replace(".data[0]", ".item()").replace(".cuda()", "")
    src = src.replace("from model import gauss_reparametrize, log_prob_laplace, log_prob_gaussian, kld_std_guss", "")
    mod = types.ModuleType("ref_evaluate")
    for k in ("gauss_reparametrize", "log_prob_laplace", "log_prob_gaussian", "kld_std_guss"):
        setattr(mod, k, getattr(M, k))
    sys.dont_write_bytecode = True
    exec(compile(src, os.path.join(lr.REF_DIR, "evaluate.py"), "exec"), mod.__dict__)
    return mod
