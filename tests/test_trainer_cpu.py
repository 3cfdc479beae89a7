"""Host-side tests of dtg_b200.trainer: the batch iterators against the reference's own iterator source
(dataloader.py:60-149, exec'd with an integer-division shim) and the loop cadence against the oracle restatement of
train.py:185-256 (oracle/loop.py; the reference loop itself is Python 2 and cannot run here)."""
import argparse
from collections import OrderedDict

import numpy as np
import pytest
import torch

import dtg  # noqa: F401
from dtg_b200 import trainer
from oracle import live_reference as lr, loop as oloop


def _data(n, seed=0):
    g = np.random.RandomState(seed)
    return g.rand(n, 1, 2, 2).astype(np.float32), g.rand(n, 1, 2, 2).astype(np.float32) + 10.0


@pytest.mark.skipif(not lr.available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("n,bs", [(10, 4), (12, 4), (7, 7), (9, 2), (5, 8)])
def test_iterators_match_reference_source(n, bs):
    RefAligned, RefUnaligned = oloop.reference_iterators()
    A, B = _data(n)
    for cls_ref, cls_ours, kw in ((RefAligned, trainer.AlignedIterator, {}), (RefAligned, trainer.AlignedIterator, {"shuffle": True}),
                                  (RefUnaligned, trainer.UnalignedIterator, {})):
        if cls_ref is RefUnaligned and bs > n:
            continue        # the reference's tail clamp indexes negatively there; not a supported configuration
        # both draw their permutations from the global numpy RNG: run them one after the other from the same seed
        np.random.seed(123)
        ref = cls_ref(A, B, batch_size=bs, **kw)
        got_ref = [list(oloop.iterate(ref)) for _ in range(3)]      # StopIteration resets; later epochs re-permute
        np.random.seed(123)
        ours = cls_ours(A, B, batch_size=bs, **kw)
        got_ours = [list(ours) for _ in range(3)]
        assert len(ref) == len(ours) and ref.n_batches == ours.n_batches
        for er, eo in zip(got_ref, got_ours):
            assert len(er) == len(eo) == ours.n_batches
            for r, o in zip(er, eo):
                assert torch.equal(r['A'], o['A']) and torch.equal(r['B'], o['B'])


class _StubModel(object):
    """records what the loop does to it"""

    def __init__(self, opt, report_flag):
        self.opt, self.calls, self.report_flag = opt, [], report_flag
        self.k = 0

    def train_instance(self, a, b, z, use_graph=False, report=True):
        self.k += 1
        self.calls.append(("step", float(a.sum()), float(b.sum()), tuple(z.shape)))
        losses = OrderedDict([("D_A", float(self.k)), ("G_A", 0.5)])
        gn = OrderedDict([("gnorm_D_A", 2.0 * self.k)])
        if self.report_flag and not report:
            return None, OrderedDict(), None
        return (losses, OrderedDict(), gn) if self.opt.monitor_gnorm else (losses, OrderedDict())

    def supervised_train_instance(self, a, b, z, use_graph=False):
        self.calls.append(("sup", float(a.sum()), float(b.sum())))
        return OrderedDict([("S_A", 1.0)])

    def save(self, name):
        self.calls.append(("save", name))

    def update_learning_rate(self):
        self.calls.append(("lr",))


def _opt(**kw):
    o = dict(epoch_count=1, niter=2, niter_decay=2, batchSize=4, nlatent=16, print_freq=8, display_freq=12, save_epoch_freq=2,
             monitor_gnorm=True, supervised=False)
    o.update(kw)
    return argparse.Namespace(**o)


def _host_stager(it, nz, gen):
    for d in it:
        yield d['A'], d['B'], torch.zeros(d['A'].size(0), nz, 1, 1)


@pytest.mark.parametrize("monitor,supervised,n", [(True, False, 10), (False, False, 12), (True, True, 9)])
def test_loop_cadence_matches_oracle(monitor, supervised, n):
    opt = _opt(monitor_gnorm=monitor, supervised=supervised)
    A, B = _data(n, seed=3)
    SA, SB = _data(6, seed=4)

    def iters():
        np.random.seed(7)
        tr = trainer.UnalignedIterator(A, B, batch_size=opt.batchSize)
        sup = trainer.AlignedIterator(SA, SB, batch_size=opt.batchSize, shuffle=True)
        return tr, sup

    # oracle loop (reference control flow)
    tr, sup = iters()
    ref_model, ref_log = _StubModel(opt, False), []
    sup_cycle = _Cycle(sup)
    total_ref = oloop.train_loop(ref_model, opt, tr, lambda m: torch.zeros(m, opt.nlatent, 1, 1), ref_log.append, sup_cycle)
    # product loop
    tr, sup = iters()
    our_model, our_log, shown = _StubModel(opt, True), [], []
    total_ours, hist = trainer.train_epochs(our_model, opt, tr, out_f=None, sup_train_dataset=_Cycle(sup), use_graph=False,
                                            log=lambda f, m: our_log.append(m), stager=_host_stager,
                                            on_display=lambda m, e, i, a, v: shown.append((e, i)))
    assert total_ours == total_ref
    assert our_model.calls == ref_model.calls           # same batches in the same order, same saves and LR decays
    ref_prints = [r for r in ref_log if r[0] == "print"]
    assert [(h[0], h[1]) for h in hist] == [(r[1], r[2]) for r in ref_prints]
    for h, r in zip(hist, ref_prints):
        assert dict(h[2]) == r[3]                        # the losses that get printed are those of the same step
    assert sum(1 for m in our_log if m.startswith("saving the model")) == sum(1 for r in ref_log if r[0] == "save")
    # display cadence: total_steps % display_freq == 0 (train.py:218)
    steps_per_epoch = tr.n_batches
    expect = [(e, i + 1) for e in range(1, 5) for i in range(steps_per_epoch)
              if (((e - 1) * steps_per_epoch + i + 1) * opt.batchSize) % opt.display_freq == 0]
    assert shown == expect


class _Cycle(object):
    """itertools.cycle over a self-resetting iterator with the reference's ``.next()`` (train.py:150-151, 212)"""

    def __init__(self, it):
        self.it = it

    def next(self):
        try:
            return next(self.it)
        except StopIteration:
            return next(self.it)

    __next__ = next


def test_format_log_matches_reference_format():
    msg = trainer.format_log(3, 40, OrderedDict([("D_A", 0.12345), ("G_A", 1.0)]), 0.0123)
    assert msg == "(epoch: 3, iters: 40, time: 0.012) D_A: 0.123 G_A: 1.000 "        # train.py:39-45
    cont = trainer.format_log(3, 40, OrderedDict([("x", 2.0)]), 0.0123, prefix=False)
    assert cont == " " * len("(epoch: 3, iters: 40, time: 0.012) ") + "x: 2.000 "


@pytest.mark.skipif(not lr.available(), reason="/root/reference not present (GPU box)")
def test_evaluate_port_matches_live_reference():
    """dtg_b200.evaluate (eval_mse_A, variational_ubo) against the reference's own evaluate.py (text shims only, see
    oracle.loop.reference_evaluate), both driving the LIVE reference model on CPU with the same RNG stream."""
    import copy
    import warnings
    from dtg_b200 import evaluate as ev
    from oracle import nets, step
    warnings.simplefilter("ignore")
    ref_ev = oloop.reference_evaluate()
    opt = step.default_opt()
    state = nets.init_model_state(seed=5, perturb=0.03)
    a, b, _ = step.synthetic_batch(3, seed=9)

    def fresh():
        m = lr.build_reference_model(copy.deepcopy(opt), state)
        m.eval()          # BatchNorm running statistics: predict_enc_params must not move them between the two runs
        return m

    torch.manual_seed(77)
    r_ubo, r_kld, r_bpp = ref_ev.variational_ubo(fresh(), a, b, 4, use_gpu=False)
    torch.manual_seed(77)
    o_ubo, o_kld, o_bpp = ev.variational_ubo(fresh(), a, b, 4)
    assert abs(o_ubo - r_ubo) <= 1e-4 * abs(r_ubo) and abs(o_kld - r_kld) <= 1e-4 * abs(r_kld) + 1e-6
    assert abs(o_bpp - r_bpp) <= 1e-4 * abs(r_bpp)
    data = [{'A': a, 'B': b}, {'A': a.flip(0), 'B': b.flip(0)}]
    assert abs(ev.eval_mse_A(data, fresh(), device="cpu") - ref_ev.eval_mse_A(data, fresh(), use_gpu=False)) < 1e-6


def test_preprocess_fields_follows_the_loader_arithmetic():
    """dataloader.py:17-34 restated with plain loops on a small field stack with NaNs (ocean cells), a constant
    channel and a fourth channel that must be dropped"""
    g = np.random.RandomState(1)
    arr = g.randn(5, 6, 7, 4).astype(np.float64) * 3 + 2
    arr[0, 1:3, 2:5, 0] = np.nan
    arr[2, :, :, 1] = 4.25                       # constant field: (x - min) / 0 -> NaN -> 0
    out = trainer.preprocess_fields(arr)
    assert out.shape == (5, 3, 6, 7) and out.dtype == np.float32
    for b in range(5):
        for c in range(3):
            f = np.nan_to_num(arr[b, :, :, c])
            lo, hi = f.min(), f.max()
            exp = np.zeros_like(f) if hi == lo else -1 + 2 * (f - lo) / (hi - lo)
            assert np.allclose(out[b, c], exp, atol=1e-6), (b, c)
    assert out.min() >= -1 - 1e-6 and out.max() <= 1 + 1e-6 and np.all(out[2, 1] == 0)
    small = trainer.preprocess_fields(arr, grid_size=4)
    assert small.shape == (5, 3, 4, 4) and np.isfinite(small).all() and abs(small).max() <= 1 + 1e-5
    same = trainer.preprocess_fields(arr[:, :6, :6], grid_size=6)          # already at grid_size: no resampling
    assert np.array_equal(same, trainer.preprocess_fields(arr[:, :6, :6]))


def test_preprocess_fields_host_matches_oracle():
    """host numpy path (vectorised) against the oracle's loop restatement of dataloader.py:17-34 + skimage<=0.14 resize
    (oracle/fields.py): shrink, enlarge (border samples mix with cval = 0), non-square input, float32 and float64"""
    from oracle import fields as ofields
    g = np.random.RandomState(3)
    arr = g.randn(4, 10, 7, 5) * 2 - 1
    arr[1, 2:4, 1:3, 2] = np.nan
    arr[3, :, :, 0] = -7.5
    for dt in (np.float64, np.float32):
        for gs in (None, 5, 16, 7):
            a = trainer.preprocess_fields(arr.astype(dt), gs)
            b = ofields.preprocess_fields(arr.astype(dt), gs)
            assert a.shape == b.shape and a.dtype == np.float32
            assert np.array_equal(a, b), (dt, gs, np.abs(a - b).max())
    # 3-D input [b, h, w]: the reference first keeps [..., :3] (the first three COLUMNS here) and then inserts the channel
    # axis at position 2 (dataloader.py:17, 20-21), so it is read as [b, h, 1, 3] -> NCHW [b, 3, h, 1]
    a3 = trainer.preprocess_fields(arr[..., 0])
    assert a3.shape == (4, 3, 10, 1) and np.array_equal(a3, ofields.preprocess_fields(arr[..., 0]))


def test_load_numpy_data_split_and_python2_shuffle(tmp_path):
    n = 230
    base = np.arange(n, dtype=np.float64)[:, None, None, None] + np.linspace(0, 1, 4 * 4 * 3).reshape(1, 4, 4, 3)
    for name, off in (("trainA", 0.0), ("trainB", 1000.0), ("testA", 0.0), ("testB", 0.0)):
        np.savez(str(tmp_path / (name + ".npz")), data=(base + off) if name.startswith("train") else base[:7])
    trA, trB, dvA, dvB, teA, teB = trainer.load_numpy_data(str(tmp_path))
    assert trA.shape == (30, 3, 4, 4) and dvA.shape == (200, 3, 4, 4) and teA.shape == (7, 3, 4, 4) and trB.shape == trA.shape
    # the permutation is Python 2.7's random.shuffle under seed 123: j = int(random() * (i + 1)) on the MT19937 stream
    import random
    rng = random.Random(123)
    idx = list(range(n))
    for i in reversed(range(1, n)):
        j = int(rng.random() * (i + 1))
        idx[i], idx[j] = idx[j], idx[i]
    plain = trainer.load_numpy_data(str(tmp_path), shuffle=False)
    allA = np.concatenate([plain[2], plain[0]])                 # unshuffled order: dev (first 200) then train
    assert np.array_equal(np.concatenate([dvA, trA]), allA[idx])
    assert np.array_equal(np.concatenate([dvB, trB]), np.concatenate([plain[3], plain[1]])[idx])   # A and B stay paired
