"""ctypes binding of libdtg_b200.so (C ABI in include/dtg_b200.h).  Fails loudly if absent."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdtg_b200.so")

BF16, F32 = 0, 1
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH = 0, 1, 2, 3
NORM_NONE, NORM_INSTANCE, NORM_COND_INSTANCE, NORM_BATCH = 0, 1, 2, 3
CONV_FWD, CONV_DGRAD = 0, 1


class Plane(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32),
                ("halo", C.c_int32), ("dtype", C.c_int32)]


class ConvArgs(C.Structure):
    _fields_ = [("mode", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32), ("stride", C.c_int32),
                ("pad", C.c_int32), ("ring", C.c_int32), ("act", C.c_int32), ("cout", C.c_int32),
                ("out_nchw_f32", C.c_int32), ("out_reflect", C.c_int32), ("out_h", C.c_int32), ("out_w", C.c_int32),
                ("fold_w", C.c_int32)]


class WgradArgs(C.Structure):
    _fields_ = [("kh", C.c_int32), ("kw", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32),
                ("pa", C.c_int32), ("qb", C.c_int32), ("fold", C.c_int32)]


class NormArgs(C.Structure):
    _fields_ = [("mode", C.c_int32), ("act", C.c_int32), ("eps", C.c_float), ("momentum", C.c_float),
                ("phase", C.c_int32), ("world_size", C.c_int32)]


class PackItem(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("rows", C.c_int32), ("rows_p", C.c_int32),
                ("cols", C.c_int32), ("cols_p", C.c_int32), ("taps", C.c_int32), ("srs", C.c_int32),
                ("scs", C.c_int32), ("dtype", C.c_int32), ("fold_kw", C.c_int32), ("fold_flip", C.c_int32),
                ("fold_fc", C.c_int32), ("s2d_k", C.c_int32)]


class LossSeg(C.Structure):
    _fields_ = [("kind", C.c_int32), ("a", C.c_void_p), ("b", C.c_void_p), ("n", C.c_int32), ("c", C.c_int32),
                ("h", C.c_int32), ("w", C.c_int32), ("target", C.c_float), ("grad_scale", C.c_float),
                ("tanh_bwd", C.c_int32), ("slot_loss", C.c_int32), ("slot_aux", C.c_int32), ("grad", C.POINTER(Plane))]


LOSS_LSGAN, LOSS_L1 = 0, 1

_lib = None

_P = C.c_void_p
_SIGS = {
    "dtg_version": (C.c_int, []),
    "dtg_launch_count": (C.c_ulonglong, []),
    "dtg_last_error": (C.c_int, [C.c_char_p, C.c_size_t]),
    "dtg_set_option": (C.c_int, [C.c_char_p, C.c_int]),
    "dtg_pack_weights": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "dtg_conv": (C.c_int, [C.POINTER(ConvArgs), C.POINTER(Plane), _P, C.c_int, C.c_int, _P, C.POINTER(Plane), _P, _P]),
    "dtg_conv_wgrad_workspace_bytes": (C.c_size_t, [C.POINTER(WgradArgs), C.POINTER(Plane), C.POINTER(Plane)]),
    "dtg_conv_wgrad": (C.c_int, [C.POINTER(WgradArgs), C.POINTER(Plane), C.POINTER(Plane), _P, _P, C.c_size_t, _P]),
    "dtg_head1_fwd": (C.c_int, [C.POINTER(Plane), _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int, C.c_int, _P]),
    "dtg_head1_dgrad": (C.c_int, [C.POINTER(Plane), _P, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Plane), _P]),
    "dtg_head1_wgrad_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "dtg_head1_wgrad": (C.c_int, [C.POINTER(Plane), C.POINTER(Plane), _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_size_t, _P]),
    "dtg_norm_workspace_bytes": (C.c_size_t, [C.POINTER(Plane)]),
    "dtg_norm_fwd": (C.c_int, [C.POINTER(NormArgs), C.POINTER(Plane), C.POINTER(Plane), _P, _P, _P, _P, _P, _P,
                               C.POINTER(Plane), _P]),
    "dtg_norm_bwd": (C.c_int, [C.POINTER(NormArgs), C.POINTER(Plane), C.POINTER(Plane), C.POINTER(Plane),
                               C.POINTER(Plane), _P, _P, _P, _P, _P, _P, C.POINTER(Plane), C.POINTER(Plane), _P]),
    "dtg_cin_affine_fwd": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "dtg_cin_affine_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P]),
    "dtg_pack_nchw": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Plane), C.c_int, C.c_int, _P]),
    "dtg_unpack_nchw": (C.c_int, [C.POINTER(Plane), C.c_int, C.c_int, _P, _P]),
    "dtg_pack_nchw_s2d": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Plane), C.c_int, C.c_int, _P]),
    "dtg_s2d_unfold_add": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "dtg_grad_gather": (C.c_int, [C.POINTER(C.POINTER(Plane)), C.POINTER(C.c_int), C.c_int, _P, _P, C.c_int,
                                  C.POINTER(Plane), _P, _P]),
    "dtg_channel_sum": (C.c_int, [C.POINTER(Plane), C.c_int, _P, _P, _P]),
    "dtg_loss_lsgan": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, _P, C.c_int, C.c_int,
                                 C.POINTER(Plane), _P, _P]),
    "dtg_loss_l1": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, _P, C.c_int, C.c_int,
                              C.POINTER(Plane), _P, _P]),
    "dtg_loss_fused": (C.c_int, [C.POINTER(LossSeg), C.c_int, _P, _P, _P]),
    "dtg_ubo_laplace": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int, C.POINTER(Plane), _P, _P]),
    "dtg_ubo_latent_step": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, _P, _P,
                                      C.c_int, _P]),
    "dtg_preprocess_fields": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "dtg_grad_sumsq": (C.c_int, [_P, C.c_size_t, C.c_float, _P, _P, _P]),
    "dtg_adam_clip": (C.c_int, [_P, _P, _P, _P, C.c_size_t, _P, _P, _P, C.c_float, _P]),
    "dtg_step_increment": (C.c_int, [_P, _P]),
}


def exported_symbols():
    return sorted(_SIGS)


def lib():
    """The loaded CDLL.  Raises RuntimeError when the extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "dtg_b200: %s is missing -- build it with `python domain-transfer-gan_b200/build.py` "
                "(there is no CPU or eager fallback)" % LIB_PATH)
        import torch  # noqa: F401  (loads libcudart before our library resolves it)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)   # AttributeError here = the .so is stale / incomplete: rebuild
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def set_option(key, value):
    """dtg_set_option: returns the previous value"""
    prev = lib().dtg_set_option(key.encode(), int(value))
    if prev < 0:
        raise RuntimeError("dtg_b200 set_option(%r) failed: %s" % (key, last_error()))
    return prev


def last_error():
    buf = C.create_string_buffer(512)
    lib().dtg_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError("dtg_b200 %s failed (%d): %s" % (what, rc, last_error()))
