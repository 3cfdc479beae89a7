"""Training loop and input staging around the fused step: the loop semantics of the reference's
``train_model`` hot loop (/root/reference/augmented_cyclegan/train.py:185-256, 307-313) and of its in-memory batch
iterators (dataloader.py:60-149), in Python 3, with the three things a millisecond-scale step needs:

  * ``StagedBatches``: pinned, double-buffered host staging; the host->device copies of batch k+1 run on a copy stream
    while step k computes (the reference issues three synchronous pageable ``.cuda()`` copies per step, train.py:198-201);
  * ``prior_z_B`` is drawn on the device (train.py:190 draws it on the host and copies it);
  * the loss dictionaries are read back only when they are printed (``print_freq``): every other step runs with
    ``report=False`` and costs no device->host synchronisation (the reference syncs >= 23 times per step, model.py:518-536).

Visualisation (PNG dumps, train.py:47-94) and the per-epoch evaluation (evaluate.py) stay with the caller: ``train_epochs``
takes them as hooks.  Nothing here touches the arithmetic of the step.
"""
import queue
import threading
import time

import numpy as np
import torch


# ---- batch iterators (dataloader.py:60-149) ------------------------------------------------------------------------
class AlignedIterator(object):
    """Iterate two arrays IN THE SAME ORDER and return dicts of minibatches (dataloader.py:60-110)."""

    def __init__(self, data_A, data_B, **kwargs):
        assert data_A.shape[0] == data_B.shape[0], 'passed data differ in number!'
        self.data_A, self.data_B = data_A, data_B
        self.num_samples = data_A.shape[0]
        self.batch_size = kwargs.get('batch_size', 100)
        self.shuffle = kwargs.get('shuffle', False)
        self.n_batches = self.num_samples // self.batch_size          # Python-2 integer division in the reference
        if self.num_samples % self.batch_size != 0:
            self.n_batches += 1
        self.reset()

    def __iter__(self):
        return self

    def reset(self):
        if self.shuffle:
            self.data_indices = np.random.permutation(self.num_samples)
        else:
            self.data_indices = np.arange(self.num_samples)
        self.batch_idx = 0

    def indices(self):
        """index arrays of the next batch (what ``next`` gathers), or None at the end of the epoch"""
        if self.batch_idx == self.n_batches:
            self.reset()
            return None
        idx = self.batch_idx * self.batch_size
        chosen = self.data_indices[idx:idx + self.batch_size]
        self.batch_idx += 1
        return chosen, chosen

    def __next__(self):
        ix = self.indices()
        if ix is None:
            raise StopIteration
        return {'A': torch.from_numpy(self.data_A[ix[0]]), 'B': torch.from_numpy(self.data_B[ix[1]])}

    next = __next__

    def __len__(self):
        return self.num_samples


class UnalignedIterator(AlignedIterator):
    """Iterate two arrays IN DIFFERENT ORDER (two independent permutations per epoch); the last batch of an epoch is
    moved back so that it is full (dataloader.py:112-152)."""

    def __init__(self, data_A, data_B, **kwargs):
        kwargs = dict(kwargs)
        kwargs.pop('shuffle', None)
        super(UnalignedIterator, self).__init__(data_A, data_B, **kwargs)

    def reset(self):
        self.data_indices = [np.random.permutation(self.num_samples) for _ in range(2)]
        self.batch_idx = 0

    def indices(self):
        if self.batch_idx == self.n_batches:
            self.reset()
            return None
        idx = self.batch_idx * self.batch_size
        if idx + self.batch_size >= len(self.data_indices[0]):
            idx = len(self.data_indices[0]) - self.batch_size
        a = self.data_indices[0][idx:idx + self.batch_size]
        b = self.data_indices[1][idx:idx + self.batch_size]
        self.batch_idx += 1
        return a, b


# ---- .npz ingestion (dataloader.py:13-59) ------------------------------------------------------------------------------
DEV_SIZE = 200


def _py2_shuffle(x, rng):
    """random.shuffle as Python 2.7 implements it (``j = int(random() * (i + 1))``); Python 3 draws the indices
    differently, so the reference's seeded train / dev split (dataloader.py:44-51) is only reproduced this way"""
    for i in reversed(range(1, len(x))):
        j = int(rng.random() * (i + 1))
        x[i], x[j] = x[j], x[i]


def _resize_fields_host(arr, gh, gw):
    """skimage.transform.resize as dataloader.py:30 calls it, vectorised numpy on [b, h, w, c] float64: the reference is
    Python 2, whose last scikit-image line (0.14) defaults to bilinear interpolation at input = scale * (o + 0.5) - 0.5,
    mode 'constant' (cval 0: samples outside the image contribute 0), no anti-aliasing.  skimage's clip to the image's
    range is a no-op on fields that have just been scaled to [-1, 1] (0 is inside the range) and is left out."""
    b, h, w, c = arr.shape
    arr = arr.astype(np.float64)

    def taps(n_in, n_out):
        x = (float(n_in) / n_out) * (np.arange(n_out) + 0.5) - 0.5
        i0, i1 = np.floor(x).astype(np.int64), np.ceil(x).astype(np.int64)
        return i0, i1, x - i0

    def gather(a, idx, axis, n_in):
        ok = (idx >= 0) & (idx < n_in)
        g = np.take(a, np.clip(idx, 0, n_in - 1), axis=axis)
        shape = [1] * a.ndim
        shape[axis] = len(idx)
        return g * ok.reshape(shape)

    r0, r1, dr = taps(h, gh)
    q0, q1, dq = taps(w, gw)
    dq = dq.reshape(1, 1, gw, 1)
    dr = dr.reshape(1, gh, 1, 1)
    rows0, rows1 = gather(arr, r0, 1, h), gather(arr, r1, 1, h)
    top = (1 - dq) * gather(rows0, q0, 2, w) + dq * gather(rows0, q1, 2, w)
    bot = (1 - dq) * gather(rows1, q0, 2, w) + dq * gather(rows1, q1, 2, w)
    return (1 - dr) * top + dr * bot


def preprocess_fields(arr, grid_size=None, device=None):
    """dataloader.py:17-34 for one array [b, h, w(, c)]: first three channels, NaN -> 0, per-sample / per-channel
    min-max scaling to [-1, 1] (constant fields -> 0), optional resize to grid_size x grid_size (skimage.transform.resize
    with the defaults of its Python-2 releases, see _resize_fields_host / dtg_preprocess_fields), NHWC -> NCHW float32.

    device=None: numpy on the host, like the reference.  device="cuda" (or a torch device): the stack is copied to the GPU
    once and the whole pipeline is ONE C-ABI call (dtg_preprocess_fields, csrc/fields.cu); returns a CUDA tensor."""
    arr = np.asarray(arr)[..., :3]
    if arr.ndim == 3:
        arr = np.expand_dims(arr, axis=2)        # the reference's quirk: [b, h, w] is read as [b, h, 1, w] (:20-21)
    if device is not None:
        return _preprocess_fields_device(arr, grid_size, device)
    arr = np.nan_to_num(arr)
    lo = arr.min((1, 2))[:, np.newaxis, np.newaxis]
    hi = arr.max((1, 2))[:, np.newaxis, np.newaxis]
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        arr = -1 + 2 * (arr - lo) / (hi - lo)
    arr = np.nan_to_num(arr, nan=0.0, posinf=0.0, neginf=0.0)
    if grid_size is not None and (arr.shape[1] != grid_size or arr.shape[2] != grid_size):
        arr = _resize_fields_host(arr, grid_size, grid_size)
    return np.ascontiguousarray(arr.transpose(0, 3, 1, 2)).astype('float32')


def _preprocess_fields_device(arr, grid_size, device, chunk=4096):
    from . import _lib
    import ctypes
    dev = torch.device("cuda", torch.cuda.current_device()) if str(device) == "cuda" else torch.device(device)
    if arr.dtype not in (np.float32, np.float64):
        arr = arr.astype(np.float64)
    b, h, w, c = arr.shape
    gh = gw = int(grid_size) if grid_size is not None else None
    gh, gw = (h, w) if gh is None else (gh, gw)
    out = torch.empty(b, c, gh, gw, dtype=torch.float32, device=dev)
    f64 = arr.dtype == np.float64
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        for s0 in range(0, b, chunk):       # bounded staging: a chunk of the stack in HBM at a time
            src = torch.from_numpy(np.ascontiguousarray(arr[s0:s0 + chunk])).to(dev)
            nb = src.shape[0]
            lohi = torch.empty(nb * c * 2, dtype=src.dtype, device=dev)
            _lib.check(_lib.lib().dtg_preprocess_fields(ctypes.c_void_p(src.data_ptr()), 1 if f64 else 0, nb, h, w, c, c, gh, gw,
                                                        ctypes.c_void_p(out[s0:s0 + nb].data_ptr()),
                                                        ctypes.c_void_p(lohi.data_ptr()), ctypes.c_void_p(stream)),
                       "preprocess_fields")
            torch.cuda.current_stream().synchronize()      # src / lohi die with this iteration
    return out


def load_numpy_data(root, shuffle=True, grid_size=None, device=None):
    """dataloader.py:13-59: {train,test}{A,B}.npz (key 'data') -> (trainA, trainB, devA, devB, testA, testB); the first
    DEV_SIZE shuffled training samples become the dev split.  Note the reference quirk kept here: a 3-D array
    [b, h, w] gets its channel axis inserted at position 2 (dataloader.py:20-21), i.e. it is read as [b, h, 1, w].
    device: preprocess on that GPU (dtg_preprocess_fields); the arrays still come back as host numpy, which is what the
    iterators of dataloader.py:112-155 index."""
    import os
    import random

    def _load(fname):
        out = preprocess_fields(np.load(os.path.join(root, fname))['data'], grid_size, device)
        return out.cpu().numpy() if device is not None else out

    trainA, trainB = _load("trainA.npz"), _load("trainB.npz")
    testA, testB = _load("testA.npz"), _load("testB.npz")
    if shuffle:
        indx = list(range(len(trainA)))
        _py2_shuffle(indx, random.Random(123))          # random.seed(123); random.shuffle(indx) under Python 2
        trainA, trainB = trainA[indx], trainB[indx]
    devA, devB = trainA[:DEV_SIZE], trainB[:DEV_SIZE]
    return trainA[DEV_SIZE:], trainB[DEV_SIZE:], devA, devB, testA, testB


# ---- input staging ---------------------------------------------------------------------------------------------------
class StagedBatches(object):
    """Wrap a batch iterator (dicts with float32 'A' / 'B' host tensors).  A staging thread copies each batch into one
    of `depth` pinned buffers and from there to the device on a copy stream, `depth - 1` batches ahead of the consumer,
    so neither the pinned memcpy nor the host->device transfer sits between two steps.  Yields (real_A, real_B,
    prior_z_B) CUDA tensors that are valid until the following ``next()``; ``prior_z_B`` ~ N(0, 1) [N, nlatent, 1, 1]
    is drawn on the device from ``generator``."""

    _END = object()

    def __init__(self, it, nlatent, device=None, generator=None, depth=2):
        if not torch.cuda.is_available():
            raise RuntimeError("dtg_b200: StagedBatches needs a CUDA device; there is no CPU fallback")
        self.it = iter(it)
        self.nlatent = nlatent
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.gen = generator
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.h2d_bytes = 0
        self._last = None
        self._free = queue.Queue()
        self._ready = queue.Queue()
        for _ in range(max(2, depth)):
            self._free.put(dict(pin_A=None, pin_B=None, dev_A=None, dev_B=None, ready=None, consumed=None))
        self._thread = threading.Thread(target=self._stage, name="dtg-stage", daemon=True)
        self._thread.start()

    def _stage(self):
        try:
            torch.cuda.set_device(self.device)
            while True:
                s = self._free.get()
                if s is None:                          # close()
                    return
                try:
                    d = next(self.it)
                except StopIteration:
                    self._ready.put(self._END)
                    return
                a, b = d['A'].float(), d['B'].float()
                if s["pin_A"] is None or s["pin_A"].shape != a.shape or s["pin_B"].shape != b.shape:
                    s.update(pin_A=torch.empty(a.shape, dtype=torch.float32).pin_memory(),
                             pin_B=torch.empty(b.shape, dtype=torch.float32).pin_memory(),
                             dev_A=torch.empty(a.shape, dtype=torch.float32, device=self.device),
                             dev_B=torch.empty(b.shape, dtype=torch.float32, device=self.device), ready=None)
                if s["ready"] is not None:
                    s["ready"].synchronize()           # the previous H2D out of this pinned buffer has completed
                s["pin_A"].copy_(a)
                s["pin_B"].copy_(b)
                if s["consumed"] is not None:
                    self.copy_stream.wait_event(s["consumed"])     # the step that read these device buffers is done
                with torch.cuda.stream(self.copy_stream):
                    s["dev_A"].copy_(s["pin_A"], non_blocking=True)
                    s["dev_B"].copy_(s["pin_B"], non_blocking=True)
                    s["ready"] = torch.cuda.Event()
                    s["ready"].record(self.copy_stream)
                self.h2d_bytes += (a.numel() + b.numel()) * 4
                self._ready.put(s)
        except BaseException as exc:                   # surface loader errors in the consumer
            self._ready.put(exc)

    def __iter__(self):
        return self

    def __next__(self):
        cur = torch.cuda.current_stream()
        if self._last is not None:
            # everything the caller issued on the batch handed out last time precedes this point in stream order
            self._last["consumed"] = torch.cuda.Event()
            self._last["consumed"].record(cur)
            self._free.put(self._last)
            self._last = None
        s = self._ready.get()
        if s is self._END:
            self._ready.put(self._END)
            raise StopIteration
        if isinstance(s, BaseException):
            raise s
        cur.wait_event(s["ready"])
        z = torch.randn(s["dev_A"].shape[0], self.nlatent, 1, 1, device=self.device, dtype=torch.float32, generator=self.gen)
        self._last = s
        return s["dev_A"], s["dev_B"], z

    next = __next__

    def close(self):
        self._free.put(None)


# ---- logging (train.py:34-45) ------------------------------------------------------------------------------------------
def print_log(out_f, message):
    if out_f is not None:
        out_f.write(message + "\n")
        out_f.flush()
    print(message)


def format_log(epoch, i, errors, t, prefix=True):
    message = '(epoch: %d, iters: %d, time: %.3f) ' % (epoch, i, t)
    if not prefix:
        message = ' ' * len(message)
    for k, v in errors.items():
        message += '%s: %.3f ' % (k, v)
    return message


# ---- the loop (train.py:185-256, 307-313) ------------------------------------------------------------------------------
def train_epochs(model, opt, train_dataset, out_f=None, sup_train_dataset=None, use_graph=True, generator=None,
                 on_display=None, on_epoch_end=None, log=print_log, stager=None):
    """Runs epochs ``opt.epoch_count .. opt.niter + opt.niter_decay`` of ``model.train_instance`` over ``train_dataset``
    (an iterator of {'A', 'B'} host batches that resets itself at StopIteration, like the reference's).  Returns
    (total_steps, history) where history lists the printed (epoch, epoch_iter, losses[, sup_losses][, gnorms]) records.

    Hooks: on_display(model, epoch, epoch_iter, real_A, visuals) at ``display_freq`` (train.py:218-241),
    on_epoch_end(model, epoch, total_steps) after the checkpoint cadence (evaluation, train.py:259-305).
    stager(iterator, nlatent, generator) -> iterator of (real_A, real_B, prior_z_B); default StagedBatches."""
    if stager is None:
        stager = lambda it, nz, gen: StagedBatches(it, nz, generator=gen)
    total_steps = 0
    history = []
    print_start_time = time.time()
    for epoch in range(opt.epoch_count, opt.niter + opt.niter_decay + 1):
        epoch_start_time = time.time()
        epoch_iter = 0
        staged = stager(_same_size_only(train_dataset), opt.nlatent, generator)
        sup_losses = None
        for real_A, real_B, prior_z_B in staged:
            total_steps += opt.batchSize                  # the reference counts batchSize even for a short batch
            epoch_iter += opt.batchSize
            show = total_steps % opt.display_freq == 0
            say = total_steps % opt.print_freq == 0
            out = model.train_instance(real_A, real_B, prior_z_B, use_graph=use_graph, report=say)
            losses, visuals = out[0], out[1]
            gnorms = out[2] if len(out) > 2 else None
            if getattr(opt, "supervised", False):
                sup = next(sup_train_dataset)
                sup_losses = model.supervised_train_instance(sup['A'].float().to(real_A.device, non_blocking=True),
                                                             sup['B'].float().to(real_A.device, non_blocking=True),
                                                             prior_z_B)
            if show and on_display is not None:
                on_display(model, epoch, epoch_iter // opt.batchSize, real_A, visuals)
            if say:
                t = (time.time() - print_start_time) / opt.batchSize
                log(out_f, format_log(epoch, epoch_iter, losses, t))
                rec = [epoch, epoch_iter, losses]
                if getattr(opt, "supervised", False):
                    log(out_f, format_log(epoch, epoch_iter, sup_losses, t, prefix=False))
                    rec.append(sup_losses)
                if opt.monitor_gnorm:
                    log(out_f, format_log(epoch, epoch_iter, gnorms, t, prefix=False) + "\n")
                    rec.append(gnorms)
                history.append(tuple(rec))
                print_start_time = time.time()
        if epoch % opt.save_epoch_freq == 0:
            log(out_f, 'saving the model at the end of epoch %d, iters %d' % (epoch, total_steps))
            model.save('latest')
        if on_epoch_end is not None:
            on_epoch_end(model, epoch, total_steps)
        log(out_f, 'End of epoch %d / %d \t Time Taken: %d sec' % (epoch, opt.niter + opt.niter_decay,
                                                                   time.time() - epoch_start_time))
        if epoch > opt.niter:
            model.update_learning_rate()
    return total_steps, history


def _same_size_only(it):
    """train.py:188-189: batches whose A and B sizes differ are skipped (without counting a step)"""
    for d in it:
        if d['A'].size(0) != d['B'].size(0):
            continue
        yield d
