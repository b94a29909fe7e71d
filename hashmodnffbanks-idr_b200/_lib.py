"""ctypes binding of libidrk.so (include/idrk.h).  Fails loudly: no fallback of any kind."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libidrk.so")

MAX_LEVELS = 32
HASH_REFERENCE = 0
HASH_TRILINEAR = 1


class IdrkError(RuntimeError):
    pass


class HashGridDesc(ctypes.Structure):
    """Mirror of idrk_hashgrid_t."""
    _fields_ = [
        ("n_levels", ctypes.c_int32),
        ("n_feat", ctypes.c_int32),
        ("frac_mode", ctypes.c_int32),
        ("n_fourier", ctypes.c_int32),
        ("res", ctypes.c_float * MAX_LEVELS),
        ("rows", ctypes.c_uint32 * MAX_LEVELS),
        ("tables", ctypes.c_void_p * MAX_LEVELS),
        ("fourier_B", ctypes.c_void_p),
    ]


_lib = None

# every symbol include/idrk.h declares (tests check the library exports all of them)
EXPORTS = ["idrk_version", "idrk_device_sm_count", "idrk_hash_encode_fwd", "idrk_hash_encode_bwd",
           "idrk_posenc_fwd", "idrk_posenc_bwd"]

_ARG_ERRORS = {-1: "bad argument", -2: "pointer or leading dimension not 16-byte aligned",
               -3: "unsupported configuration", -4: "CUDA driver entry point unavailable"}


def lib():
    """Loads libidrk.so on first use; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise IdrkError("libidrk.so is missing (%s). Build it with `python hashmodnffbanks-idr_b200/csrc/build.py`; "
                            "there is no CPU or PyTorch fallback." % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def _declare(L):
    c = ctypes
    vp, i32, i64, f32 = c.c_void_p, c.c_int32, c.c_int64, c.c_float
    L.idrk_version.restype = c.c_int
    L.idrk_device_sm_count.argtypes = [c.POINTER(c.c_int)]
    L.idrk_hash_encode_fwd.argtypes = [c.POINTER(HashGridDesc), vp, i64, i32, vp, i32, vp, vp]
    L.idrk_hash_encode_bwd.argtypes = [c.POINTER(HashGridDesc), vp, i64, i32, vp, i32, c.POINTER(vp), vp, vp]
    fp = c.POINTER(c.c_float)
    L.idrk_posenc_fwd.argtypes = [vp, i64, i32, i32, fp, i32, i32, vp, i32, vp]
    L.idrk_posenc_bwd.argtypes = [vp, i64, i32, i32, fp, i32, i32, vp, i32, vp, i32, vp]
    for fn in EXPORTS:
        getattr(L, fn).restype = c.c_int


def check(rc, what):
    if rc == 0:
        return
    if rc < 0:
        raise IdrkError("%s: %s (code %d)" % (what, _ARG_ERRORS.get(rc, "error"), rc))
    raise IdrkError("%s: CUDA error %d" % (what, rc))


def require_cuda(t: torch.Tensor, name="tensor"):
    if not t.is_cuda:
        raise IdrkError("%s must live on a CUDA device: idrk has no CPU path" % name)


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)
