"""Thin Python wrappers over the C ABI: plane tensors, weight packing, conv / wgrad / norm calls.

Everything here only marshals pointers; all arithmetic happens in libdtg_b200.so on the current
CUDA stream.  Nothing falls back to torch ops.
"""
import ctypes as C
import os

import torch

from . import _lib as L

_DT = {torch.bfloat16: L.BF16, torch.float32: L.F32}


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# ---- lanes: independent branches of a step on parallel CUDA streams ------------------------------------------------
_LANE = 0


def lane():
    """index of the lane (stream) the caller is issuing on; scratch buffers are per lane"""
    return _LANE


class Lanes(object):
    """Fork / join of independent branches of one step onto lane CUDA streams.  begin() forks the lanes off the
    caller's current stream (under torch.cuda.graph: the capturing stream, so the lanes join the capture through the
    fork event and the branches become parallel paths of the SAME CUDA graph); end() joins them back.  Tasks on one
    lane run in issue order; cross-lane dependencies are the events returned by run().  With enabled=False every task runs on lane 0 in
    issue order (the issue order must therefore be a valid serial schedule)."""

    def __init__(self, n_lanes, device, chain_priority=0, comp_priority=0):
        # DTG_LANE_PRIO="chain,companion" overrides the stream priorities (measured on B200: equal priorities are best;
        # favouring the chains pushes the weight gradients into a tail, favouring the companions delays the chains)
        if os.environ.get("DTG_LANE_PRIO"):
            chain_priority, comp_priority = [int(v) for v in os.environ["DTG_LANE_PRIO"].split(",")]
        # lanes carry the dependency chains (forward / norm-backward / dgrad), companions the off-chain work
        self.lane_streams = [torch.cuda.Stream(device=device, priority=chain_priority) for _ in range(n_lanes)]
        self.comp = [torch.cuda.Stream(device=device, priority=comp_priority) for _ in range(n_lanes)]   # companion of each lane
        self.enabled = True
        self.companions = True
        self.streams = None
        self.main = None

    def begin(self):
        global ACTIVE
        self.main = torch.cuda.current_stream()
        if not self.enabled:
            self.streams = [self.main]
            ACTIVE = None
            return
        self.streams = self.lane_streams
        ACTIVE = self
        ev = torch.cuda.Event()
        ev.record(self.main)
        for s in self.streams:
            s.wait_event(ev)

    def run(self, lane_idx, fn, after=()):
        """issue fn() on lane `lane_idx` after the events in `after`; returns the completion event"""
        global _LANE
        if not self.enabled:
            fn()
            return None
        s = self.streams[lane_idx]
        for ev in after:
            if ev is not None:
                s.wait_event(ev)
        prev, _LANE = _LANE, lane_idx
        try:
            with torch.cuda.stream(s):
                fn()
        finally:
            _LANE = prev
        ev = torch.cuda.Event()
        ev.record(s)
        return ev

    def end(self):
        global ACTIVE
        if self.enabled:
            for s in self.streams:
                self.main.wait_stream(s)
        self.streams = None
        ACTIVE = None

    # A lane's companion stream carries work that hangs off the lane's chain without feeding it back (the weight
    # gradients of a backward pass: each depends on one node of the norm-backward / dgrad chain and only the
    # optimizer consumes it), so it overlaps the chain instead of lengthening it.
    def companion_run(self, fn):
        """issue fn() on the companion stream of the lane now issuing, after everything issued on the lane so far"""
        global _LANE
        if not self.companions:
            fn()
            return
        cur, cs = torch.cuda.current_stream(), self.comp[_LANE]
        cs.wait_stream(cur)
        prev, _LANE = _LANE, _LANE + len(self.comp)
        try:
            with torch.cuda.stream(cs):
                fn()
        finally:
            _LANE = prev

    def companion_join(self):
        """the lane now issuing waits for its companion stream"""
        if self.companions:
            torch.cuda.current_stream().wait_stream(self.comp[_LANE])


ACTIVE = None       # the Lanes object between begin() and end()


def off_chain(fn):
    """run fn on the current lane's companion stream when a lane schedule is active, else inline"""
    if ACTIVE is not None:
        ACTIVE.companion_run(fn)
    else:
        fn()


def off_chain_join():
    if ACTIVE is not None:
        ACTIVE.companion_join()


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def cpad(c, dtype):
    """stored channel count: multiple of 16 bytes, at least 16 channels (UMMA N granularity)."""
    return max(16, (c + 15) // 16 * 16)


def fold_channels(dtype):
    """channels of a 16-byte-per-pixel plane (the unit of the kw-folded 7x7 kernels): 8 bf16 / 4 fp32"""
    return 8 if dtype == torch.bfloat16 else 4


def cpad_small(c, dtype):
    """stored channels of network inputs / head gradients: a 16-byte pixel when the tensor fits (so that the 7x7
    head / tail can read it kw-folded), else cpad"""
    fc = fold_channels(dtype)
    return fc if c <= fc else cpad(c, dtype)


class PlaneT:
    """NHWC activation / gradient buffer with optional halo ring: tensor [n, h+2*halo, w+2*halo, c]."""

    __slots__ = ("t", "n", "h", "w", "c", "halo", "dtype", "_s", "_buf", "s2d")

    def __init__(self, n, h, w, c, halo=0, dtype=torch.bfloat16, device="cuda", s2d=0):
        self.n, self.h, self.w, self.c, self.halo, self.dtype = n, h, w, c, halo, dtype
        # s2d = cp > 0: space-to-depth plane of a [n, 2h, 2w, cp] image (c == 4*cp); pack_nchw fills it accordingly
        self.s2d = s2d
        # 256 bytes of zero slack behind the plane: kw-folded TMA views read one 128-byte row past the last pixel
        numel = n * (h + 2 * halo) * (w + 2 * halo) * c
        self._buf = torch.zeros(numel + 256 // torch.empty((), dtype=dtype).element_size(), dtype=dtype, device=device)
        self.t = self._buf[:numel].view(n, h + 2 * halo, w + 2 * halo, c)
        self._s = L.Plane(self.t.data_ptr(), n, h, w, c, halo, _DT[dtype])

    @property
    def s(self):
        return C.byref(self._s)

    def interior(self):
        hl = self.halo
        return self.t[:, hl:hl + self.h, hl:hl + self.w, :]

    def to_nchw(self, c=None):
        return self.interior()[..., :c].permute(0, 3, 1, 2).float().contiguous()

    def batch_slice(self, i0, i1):
        """view of images [i0, i1) as a plane of batch i1-i0 (shares memory)"""
        v = PlaneT.__new__(PlaneT)
        v.n, v.h, v.w, v.c, v.halo, v.dtype = i1 - i0, self.h, self.w, self.c, self.halo, self.dtype
        v.t = self.t[i0:i1]
        v.s2d = self.s2d
        v._buf = self._buf
        v._s = L.Plane(v.t.data_ptr(), v.n, v.h, v.w, v.c, v.halo, _DT[self.dtype])
        return v

    @staticmethod
    def from_nchw(x, halo=0, dtype=torch.bfloat16, c_store=None, reflect=True):
        """test helper: build a plane from an NCHW tensor through the library's own pack kernel."""
        n, c, h, w = x.shape
        p = PlaneT(n, h, w, c_store or cpad(c, dtype), halo, dtype, x.device)
        pack_nchw(x.float().contiguous(), p, 0, reflect=reflect)
        return p


NULL_PLANE = C.POINTER(L.Plane)()


def pack_nchw(src, dst, c_off=0, tanh_y=None, reflect=True):
    """reflect=False leaves dst's halo untouched (zero padding: planes are zero-initialised)"""
    n, c, h, w = src.shape
    assert src.dtype == torch.float32 and src.is_contiguous()
    if getattr(dst, "s2d", 0):
        assert tanh_y is None
        L.check(L.lib().dtg_pack_nchw_s2d(_ptr(src), n, c, h, w, dst.s, dst.s2d, c_off, _stream()), "pack_nchw_s2d")
        return
    L.check(L.lib().dtg_pack_nchw(_ptr(src), _ptr(tanh_y), n, c, h, w, dst.s, c_off, 1 if reflect else 0, _stream()),
            "pack_nchw")


def s2d_unfold_add(dw2, dw, cp):
    """dw [cout][cin][k][k] += the space-to-depth weight gradient dw2 [cout][4*cp][3][3]; clears dw2"""
    cout, cin, k, _ = dw.shape
    assert dw2.numel() == cout * 4 * cp * 9 and dw2.dtype == torch.float32 and dw.dtype == torch.float32
    L.check(L.lib().dtg_s2d_unfold_add(_ptr(dw2), _ptr(dw), cout, cin, k, cp, _stream()), "s2d_unfold_add")


def unpack_nchw(src, c, c_off=0, out=None):
    if out is None:
        out = torch.empty(src.n, c, src.h, src.w, dtype=torch.float32, device=src.t.device)
    L.check(L.lib().dtg_unpack_nchw(src.s, c_off, c, _ptr(out), _stream()), "unpack_nchw")
    return out


class PackTable:
    """Device-resident table of weight-pack items; one launch repacks every conv weight of a network."""

    def __init__(self, device="cuda"):
        self.items = []
        self.keep = []
        self.dev = None
        self.max_elems = 0
        self.device = device

    def add(self, src, rows, cols, taps, srs, scs, dtype, fold_kw=0, fold_flip=False, s2d_k=0, s2d_cp=0):
        """src: fp32 tensor (PyTorch layout, contiguous).  Returns the packed destination tensor
        [taps, rows_p, cols_p] with dst[t][r][c] = src.flat[(r*srs + c*scs)*taps + t].
        fold_kw = KW > 0: kw-folded packing (taps = KH): dst[kh][r][j*fc + b] = src.flat[((r*srs + b*scs)*KH + kh)*KW + kw(j)],
        128 bytes per row (8 filter-column slots of fc channels)."""
        rows_p = max(16, (rows + 15) // 16 * 16)
        q = 8 if dtype == torch.bfloat16 else 4
        if fold_kw and int(fold_flip) == 2:     # filter column in the ROWS (dtg_conv fold_w = 2): [KH][32][cols_p]
            assert fold_kw * rows <= 28
            rows_p, cols_p = 32, (cols + q - 1) // q * q
        else:
            cols_p = 8 * q if fold_kw else (cols + q - 1) // q * q
            assert not fold_kw or (cols <= q and fold_kw <= 8)
        if s2d_k:       # dtg_pack_item.s2d_k: 9 taps, 4 * cp columns
            assert taps == 9 and cols <= s2d_cp and not fold_kw
            cols_p = 4 * s2d_cp
        dst = torch.zeros(taps, rows_p, cols_p, dtype=dtype, device=self.device)
        self.items.append(L.PackItem(src.data_ptr(), dst.data_ptr(), rows, rows_p, cols, cols_p, taps, srs, scs,
                                     _DT[dtype], fold_kw, int(fold_flip), s2d_cp if s2d_k else q, s2d_k))
        self.keep.append((src, dst))
        self.max_elems = max(self.max_elems, dst.numel())
        self.dev = None
        return dst

    def run(self):
        if not self.items:
            return
        if self.dev is None:
            arr = (L.PackItem * len(self.items))(*self.items)
            raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).clone()
            self.dev = raw.to(self.device)
        L.check(L.lib().dtg_pack_weights(_ptr(self.dev), len(self.items), self.max_elems, _stream()), "pack_weights")


def pack_conv_weight(w, dtype, kind):
    """Convenience (tests / small modules): pack one weight immediately.
    kind: 'fwd'   conv weight [co,ci,kh,kw]  -> rows co, cols ci        (Conv2d forward)
          'fwd_fold'   same, kw-folded for a <= 16-byte-per-pixel input (dtg_conv fold_w)
          'dgrad' conv weight [co,ci,kh,kw]  -> rows ci, cols co        (Conv2d data gradient)
          'dgrad_fold' same, kw-folded + flipped for a small-channel dy (dtg_conv fold_w, DGRAD)
          'fwd_kwn'    conv weight [co,ci,kh,kw] -> [kh][(kw, co) padded to 32][ci]   (dtg_conv fold_w = 2)
          'dgrad_kwn'  conv weight [co,ci,kh,kw] -> [kh][(kw, ci) padded to 32][co]   (dtg_conv fold_w = 2, DGRAD)
          'tfwd'  convT weight [ci,co,kh,kw] -> rows co, cols ci        (ConvTranspose2d forward)
          'tdgrad' convT weight [ci,co,kh,kw]-> rows ci, cols co        (ConvTranspose2d data gradient)"""
    tab = PackTable(w.device)
    dst = add_packed(tab, w, dtype, kind)
    tab.run()
    return dst


def add_packed(tab, w, dtype, kind, s2d_cp=0):
    d0, d1, kh, kw = w.shape
    taps = kh * kw
    if kind == "fwd_s2d":              # stride-2 pad-1 first layer over a space-to-depth input plane
        return tab.add(w, d0, d1, 9, d1, 1, dtype, s2d_k=kh, s2d_cp=s2d_cp)
    if kind in ("fwd", "tdgrad"):      # rows = dim0, cols = dim1
        return tab.add(w, d0, d1, taps, d1, 1, dtype)
    if kind in ("dgrad", "tfwd"):      # rows = dim1, cols = dim0
        return tab.add(w, d1, d0, taps, 1, d1, dtype)
    if kind == "fwd_fold":
        return tab.add(w, d0, d1, kh, d1, 1, dtype, fold_kw=kw)
    if kind == "dgrad_fold":
        return tab.add(w, d1, d0, kh, 1, d1, dtype, fold_kw=kw, fold_flip=True)
    if kind == "fwd_kwn":              # filter column in GEMM-N (conv_tail7.cu): rows = (kw, dim0), cols = dim1, taps = kh
        return tab.add(w, d0, d1, kh, d1, 1, dtype, fold_kw=kw, fold_flip=2)
    if kind == "dgrad_kwn":            # the same for the data gradient: rows = (kw, dim1), cols = dim0
        return tab.add(w, d1, d0, kh, 1, d1, dtype, fold_kw=kw, fold_flip=2)
    raise ValueError(kind)


class OpLog:
    """Optional log of every ops.* call in issue order: (name, kernels launched, algorithmic FLOPs, algorithmic bytes).
    bench.py captures one serial step with it and lines the records up with the CUPTI kernel records of the replayed
    graph (the k-th dtg:: kernel on the stream belongs to the op whose launch range covers k), which gives exact
    per-op / per-kernel device times without host launch gaps."""

    def __init__(self):
        self.records = []


PROFILE = None    # set to an OpLog to record the calls


def conv(x, wp, bias, out, *, mode=L.CONV_FWD, kh, kw, stride=1, pad=0, ring=0, act=L.ACT_NONE, cout,
         out_h, out_w, out_reflect=False, out_nchw=None, cin=None, fold_w=False):
    """x: PlaneT; wp: packed weight [taps, rows_p, cols_p]; out: PlaneT or None with out_nchw fp32 tensor.
    cin: real input channels (FLOP accounting only)."""
    a = L.ConvArgs(mode, kh, kw, stride, pad, ring, act, cout, 1 if out_nchw is not None else 0,
                   1 if out_reflect else 0, out_h, out_w, int(fold_w))
    assert wp.shape[0] == (kh if fold_w else kh * kw)
    rc = L.lib().dtg_conv(C.byref(a), x.s, _ptr(wp), wp.shape[1], wp.shape[2], _ptr(bias),
                          out.s if out is not None else NULL_PLANE, _ptr(out_nchw), _stream())
    L.check(rc, "conv")


def flat_dgrad_eligible(n, h, w, c_dy, c_dx, ring, dtype):
    """geometry test of the flat-raster data gradient of conv_patch2.cu (dtg_conv DGRAD with a dy plane whose ZERO halo equals
    the ring): full 128-byte channel chunks on both sides, <= 128 output channels, a patch of 128 + 2 * (w + 2 ring + 1)
    pixels <= 256, two patch units per chunk plus three weight stages in 227 KB, pixel count divisible by 8"""
    es = 2 if dtype == torch.bfloat16 else 4
    wr, hr = w + 2 * ring, h + 2 * ring
    if ring < 1 or (c_dy * es) % 128 or (c_dx * es) % 128 or c_dx > 128 or (n * hr * wr) % 8:
        return False
    rows = 128 + 2 * (wr + 1)
    if rows > 256:
        return False
    unit = (rows * 128 + 1023) // 1024 * 1024
    kchunks = c_dy * es // 128
    return 2 * kchunks * unit + 3 * c_dx * 128 + 2048 + 4 * 9 * 1024 <= 227 * 1024


def tail_kwn_eligible(cin_stored, k, cout, w, dtype):
    """geometry test of dtg_conv fold_w = 2 (conv_tail7.cu: try_launch_tail7) for a k x k 'same' head on a halo-free plane
    with cin_stored channels and image width w: row bytes 32 / 64 / 128, (kw, cout) <= 28 GEMM columns, the width a divisor
    of 128, and two patch stages (128 / w + k - 1 image rows each) next to the weights and the staging tile in 227 KB"""
    rb = cin_stored * (2 if dtype == torch.bfloat16 else 4)
    if rb not in (32, 64, 128) or not (1 <= cout <= 4) or k % 2 == 0 or k > 8 or k * cout > 28:
        return False
    if w < 8 or w > 128 or 128 % w:
        return False
    stage = ((128 // w + k - 1) * w * rb + 1023) // 1024 * 1024
    fixed = 1024 + (k * 32 * rb + 1023) // 1024 * 1024 + 1024 + 2 * 128 * 29 * 4
    return 2 * stage + fixed <= 227 * 1024


_ws_cache = {}
_ws_retired = []      # outgrown scratch buffers: captured CUDA graphs may still hold their addresses, so they are never freed


def workspace(nbytes, device="cuda", tag="default"):
    """Grow-only scratch buffer per (device, tag, lane); contents are never assumed to persist.  A buffer that is outgrown
    is retired, not freed: an earlier captured graph (another batch shape) keeps writing to it on replay."""
    key = (str(device), tag, _LANE)
    t = _ws_cache.get(key)
    if t is None or t.numel() < nbytes:
        if t is not None:
            _ws_retired.append(t)
        t = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = t
    return t


def conv_wgrad(p, q, dw, *, kh, kw, stride=1, pad=0, pa, qb, ws=None, fold=0):
    """dw[a][b][kh][kw] += sum_pix p[pix][a] * q[pix*stride + tap - pad][b]; dw fp32 contiguous.
    fold: 1 = q is a 16-byte-per-pixel plane read kw-folded, 2 = p is (dtg_wgrad_args.fold)."""
    a = L.WgradArgs(kh, kw, stride, pad, pa, qb, fold)
    need = L.lib().dtg_conv_wgrad_workspace_bytes(C.byref(a), p.s, q.s)
    if need == 0:
        raise RuntimeError("dtg_b200 conv_wgrad: invalid geometry: " + L.last_error())
    if ws is None:
        ws = workspace(need, dw.device, "wgrad")
    assert dw.dtype == torch.float32 and dw.is_contiguous() and dw.numel() == pa * qb * kh * kw
    L.check(L.lib().dtg_conv_wgrad(C.byref(a), p.s, q.s, _ptr(dw), _ptr(ws), ws.numel(), _stream()), "conv_wgrad")


def head1_fwd(x, w, bias, out_nchw, pad):
    """cout == 1 head conv (dtg_head1_fwd): w is the fp32 master weight [1, cin, kh, kw]"""
    _, cin, kh, kw = w.shape
    n, _, oh, ow = out_nchw.shape
    L.check(L.lib().dtg_head1_fwd(x.s, _ptr(w), _ptr(bias), cin, kh, kw, pad, _ptr(out_nchw), oh, ow, _stream()), "head1_fwd")


def head1_dgrad(dy, w, dx, pad):
    _, cin, kh, kw = w.shape
    L.check(L.lib().dtg_head1_dgrad(dy.s, _ptr(w), cin, kh, kw, pad, dx.s, _stream()), "head1_dgrad")


def head1_wgrad(dy, x, dw, pad):
    _, cin, kh, kw = dw.shape
    need = L.lib().dtg_head1_wgrad_workspace_bytes(cin, kh, kw)
    ws = workspace(need, dw.device, "head1_wgrad")
    L.check(L.lib().dtg_head1_wgrad(dy.s, x.s, _ptr(dw), cin, kh, kw, pad, _ptr(ws), ws.numel(), _stream()), "head1_wgrad")


def norm_workspace_floats(x):
    return L.lib().dtg_norm_workspace_bytes(x.s) // 4


class NormState:
    """Per-layer fp32 side buffers of one norm invocation (saved statistics, coefficients, sums)."""

    __slots__ = ("stats", "coef", "sums", "ws")

    def __init__(self, x):
        dev = x.t.device
        self.stats = torch.zeros(x.n, x.c, 2, dtype=torch.float32, device=dev)
        self.coef = torch.zeros(x.n, x.c, 2, dtype=torch.float32, device=dev)
        self.sums = torch.zeros(x.n, x.c, 2, dtype=torch.float32, device=dev)
        self.ws = torch.zeros(norm_workspace_floats(x), dtype=torch.float32, device=dev)


def norm_fwd(x, out, st, *, mode, act, gamma=None, beta=None, residual=None, bn_running=None, eps=1e-5,
             momentum=0.1, phase=0, world_size=1):
    a = L.NormArgs(mode, act, eps, momentum, phase, world_size)
    rc = L.lib().dtg_norm_fwd(C.byref(a), x.s, residual.s if residual is not None else NULL_PLANE, _ptr(gamma),
                              _ptr(beta), _ptr(bn_running), _ptr(st.stats), _ptr(st.coef), _ptr(st.ws), out.s, _stream())
    L.check(rc, "norm_fwd")


def norm_bwd(dy, dx, st, *, mode, act, y=None, x=None, gamma=None, dy2=None, d_res=None, d_gamma=None, d_beta=None,
             want_sums=False, phase=0, world_size=1, defer_channel=False):
    """defer_channel: the per-channel d_gamma / d_beta reduction (a launch-latency-bound kernel that only the optimizer
    consumes) goes to the lane's companion stream instead of the data-gradient chain (dtg_norm_bwd phases 4 + 3)"""
    P = lambda p: p.s if p is not None else NULL_PLANE
    sums = _ptr(st.sums) if (want_sums or mode == L.NORM_COND_INSTANCE) else C.c_void_p(0)

    def call(ph):
        a = L.NormArgs(mode, act, 1e-5, 0.1, ph, world_size)
        rc = L.lib().dtg_norm_bwd(C.byref(a), dy.s, P(dy2), P(y), P(x), _ptr(st.stats), _ptr(gamma), sums,
                                  _ptr(d_gamma), _ptr(d_beta), _ptr(st.ws), dx.s, P(d_res), _stream())
        L.check(rc, "norm_bwd")

    if (defer_channel and phase == 0 and ACTIVE is not None and mode in (L.NORM_NONE, L.NORM_INSTANCE)
            and (d_gamma is not None or d_beta is not None)):
        call(4)
        off_chain(lambda: call(3))
    else:
        call(phase)


def cin_affine_fwd(z, ws, bs, wb, bb, gamma, beta):
    n, nz = z.shape[0], z.shape[1]
    c = ws.shape[0]
    L.check(L.lib().dtg_cin_affine_fwd(_ptr(z), _ptr(ws), _ptr(bs), _ptr(wb), _ptr(bb), n, c, nz, _ptr(gamma),
                                       _ptr(beta), _stream()), "cin_affine_fwd")


def cin_affine_bwd(z, ws, wb, gamma, beta, sums, d_ws, d_bs, d_wb, d_bb, d_z):
    n, nz = z.shape[0], z.shape[1]
    c = ws.shape[0]
    L.check(L.lib().dtg_cin_affine_bwd(_ptr(z), _ptr(ws), _ptr(wb), _ptr(gamma), _ptr(beta), _ptr(sums), n, c, nz,
                                       _ptr(d_ws), _ptr(d_bs), _ptr(d_wb), _ptr(d_bb), _ptr(d_z), _stream()),
            "cin_affine_bwd")


def grad_gather(srcs, c_offs, c, out=None, tanh_y=None, out_nchw=None, add_nchw=None):
    arr = (C.POINTER(L.Plane) * len(srcs))(*[C.pointer(s._s) for s in srcs])
    offs = (C.c_int * len(srcs))(*c_offs)
    if out is None:
        s0 = srcs[0]
        dummy = L.Plane(None, s0.n, s0.h, s0.w, c, 0, _DT[s0.dtype])
        outp = C.byref(dummy)
    else:
        outp = out.s
    L.check(L.lib().dtg_grad_gather(arr, offs, len(srcs), _ptr(add_nchw), _ptr(tanh_y), c, outp, _ptr(out_nchw), _stream()), "grad_gather")


_cs_ws = {}


def channel_sum(x, c, d_bias):
    dev = (str(d_bias.device), _LANE)
    if dev not in _cs_ws:
        _cs_ws[dev] = torch.zeros(2048, dtype=torch.float32, device=d_bias.device)
    L.check(L.lib().dtg_channel_sum(x.s, c, _ptr(d_bias), _ptr(_cs_ws[dev]), _stream()), "channel_sum")


def loss_lsgan(pred, target, grad_scale, scalars, slot_loss, slot_mean, dpred, ws):
    n, _, h, w = pred.shape
    L.check(L.lib().dtg_loss_lsgan(_ptr(pred), n, h, w, float(target), float(grad_scale), _ptr(scalars), slot_loss,
                                   slot_mean, dpred.s if dpred is not None else NULL_PLANE, _ptr(ws), _stream()), "loss_lsgan")


def loss_l1(a, b, grad_scale, tanh_bwd, scalars, slot_loss, slot_aux, da, ws):
    n, c, h, w = a.shape
    L.check(L.lib().dtg_loss_l1(_ptr(a), _ptr(b), n, c, h, w, float(grad_scale), 1 if tanh_bwd else 0, _ptr(scalars),
                                slot_loss, slot_aux, da.s if da is not None else NULL_PLANE, _ptr(ws), _stream()), "loss_l1")


def lsgan_seg(pred, target, grad_scale, slot_loss, slot_mean, dpred):
    """one LSGAN term of loss_fused (same meaning as loss_lsgan's arguments)"""
    n, _, h, w = pred.shape
    return L.LossSeg(L.LOSS_LSGAN, pred.data_ptr(), None, n, 1, h, w, float(target), float(grad_scale), 0, slot_loss, slot_mean,
                     C.pointer(dpred._s) if dpred is not None else None), (pred, dpred)


def l1_seg(a, b, grad_scale, tanh_bwd, slot_loss, slot_aux, da):
    n, c, h, w = a.shape
    return L.LossSeg(L.LOSS_L1, a.data_ptr(), b.data_ptr(), n, c, h, w, 0.0, float(grad_scale), 1 if tanh_bwd else 0, slot_loss,
                     slot_aux, C.pointer(da._s) if da is not None else None), (a, b, da)


def loss_fused(segs, scalars, ws):
    """several loss terms in ONE launch (dtg_loss_fused); segs: results of lsgan_seg / l1_seg; ws: nseg x 1024 floats"""
    arr = (L.LossSeg * len(segs))(*[s for s, _ in segs])
    assert ws.numel() >= 1024 * len(segs)
    L.check(L.lib().dtg_loss_fused(arr, len(segs), _ptr(scalars), _ptr(ws), _stream()), "loss_fused")


def ubo_laplace(fake, real, logvar_b, scalars, slot_logp, dfake, ws):
    """evaluate.py:93-94: mean over samples of the summed Laplace log-likelihood + seed gradient of its negative"""
    n, c, h, w = fake.shape
    assert logvar_b.numel() == c * h * w and real.shape == fake.shape
    L.check(L.lib().dtg_ubo_laplace(_ptr(fake), _ptr(real), _ptr(logvar_b), n, c, h, w, _ptr(scalars), slot_logp,
                                    dfake.s if dfake is not None else NULL_PLANE, _ptr(ws), _stream()), "ubo_laplace")


def ubo_latent_step(mu, logvar, sq_mu, sq_logvar, eps_cur, eps_next, dz, lr, alpha, rms_eps, z_out, scalars, slot_kld):
    """evaluate.py:101, 118-123: KLD of the current iterate, RMSprop step on (mu, logvar), next z"""
    n, nz = mu.shape
    L.check(L.lib().dtg_ubo_latent_step(_ptr(mu), _ptr(logvar), _ptr(sq_mu), _ptr(sq_logvar), _ptr(eps_cur), _ptr(eps_next),
                                        _ptr(dz), n, nz, float(lr), float(alpha), float(rms_eps), _ptr(z_out), _ptr(scalars),
                                        slot_kld, _stream()), "ubo_latent_step")


def grad_sumsq(g, grad_scale, out, ws):
    L.check(L.lib().dtg_grad_sumsq(_ptr(g), g.numel(), float(grad_scale), _ptr(out), _ptr(ws), _stream()), "grad_sumsq")


def adam_clip(p, g, m, v, hyper, sumsq, step_dev, grad_scale=1.0):
    L.check(L.lib().dtg_adam_clip(_ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), _ptr(hyper), _ptr(sumsq),
                                  _ptr(step_dev), float(grad_scale), _stream()), "adam_clip")


def step_increment(step_dev):
    L.check(L.lib().dtg_step_increment(_ptr(step_dev), _stream()), "step_increment")


# ---- op log (see OpLog): algorithmic work of the heavy ops, launch ranges of all of them ------------------------------
def _conv_work(x, wp, bias, out, *, mode=L.CONV_FWD, kh, kw, cout, out_h, out_w, cin=None, **_):
    # algorithmic MACs = (pixels of the low-resolution side) * cin * cout * taps, for fwd and dgrad alike
    pix = x.n * (out_h * out_w if mode == L.CONV_FWD else x.h * x.w)
    return 2.0 * pix * (cin or wp.shape[2]) * cout * kh * kw, 0.0


def _wgrad_work(p, q, dw, *, kh, kw, pa, qb, **_):
    return 2.0 * p.n * p.h * p.w * pa * qb * kh * kw, 0.0


def _plane_bytes(p):
    return float(p.n * p.h * p.w * p.c * p.t.element_size())


def _norm_fwd_work(x, out, st, *, residual=None, **_):
    # read x (+ residual), write y
    return 0.0, _plane_bytes(x) * (2 + (1 if residual is not None else 0))


def _norm_bwd_work(dy, dx, st, *, mode, act, y=None, x=None, dy2=None, d_res=None, **_):
    # read dy (+ dy2, + y for the activation mask, + x for the statistics), write dx (+ residual-branch gradient)
    passes = 2 + (1 if dy2 is not None else 0) + (1 if act != L.ACT_NONE else 0) + (1 if mode != L.NORM_NONE else 0) + \
        (1 if d_res is not None else 0)
    return 0.0, _plane_bytes(dx) * passes


def _logged(name, fn, work=None):
    def w(*a, **k):
        if PROFILE is None:
            return fn(*a, **k)
        lc0 = L.lib().dtg_launch_count()
        r = fn(*a, **k)
        fl, by = work(*a, **k) if work is not None else (0.0, 0.0)
        PROFILE.records.append((name, int(L.lib().dtg_launch_count() - lc0), fl, by))
        return r
    w.__name__, w.__doc__ = fn.__name__, fn.__doc__
    return w


for _n, _w in (("conv", _conv_work), ("conv_wgrad", _wgrad_work), ("norm_fwd", _norm_fwd_work), ("norm_bwd", _norm_bwd_work),
               ("pack_nchw", None), ("unpack_nchw", None), ("s2d_unfold_add", None), ("head1_fwd", None), ("head1_dgrad", None),
               ("head1_wgrad", None), ("cin_affine_fwd", None), ("cin_affine_bwd", None), ("grad_gather", None),
               ("channel_sum", None), ("loss_lsgan", None), ("loss_l1", None), ("loss_fused", None), ("grad_sumsq", None), ("adam_clip", None),
               ("step_increment", None)):
    globals()[_n] = _logged(_n, globals()[_n], _w)
PackTable.run = _logged("pack_weights", PackTable.run)
