"""Thin typed wrappers around the C ABI (include/idrk.h): tensors in, kernels launched on the
current CUDA stream.  No arithmetic happens here and nothing falls back to PyTorch."""
import ctypes
import math
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import HashGridDesc, check, lib, ptr, require_cuda, stream_ptr


def pad4(n: int) -> int:
    return (n + 3) // 4 * 4


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    require_cuda(t, name)
    if t.dtype != torch.float32:
        raise _lib.IdrkError("%s must be float32" % name)
    return t


def rows2d(t: torch.Tensor, name: str) -> torch.Tensor:
    """Returns a 2-D fp32 CUDA tensor whose last dim is unit-stride (copying only if it must)."""
    _f32c(t, name)
    if t.dim() != 2:
        raise _lib.IdrkError("%s must be 2-D" % name)
    if t.stride(1) != 1 or (t.shape[0] > 1 and t.stride(0) < t.shape[1]):
        t = t.contiguous()
    return t


def ld_of(t: torch.Tensor) -> int:
    """Leading dimension (floats) of a 2-D row-major tensor; a single row may use any ld."""
    return t.stride(0) if t.shape[0] > 1 else max(t.shape[1], 1)


# ---------------------------------------------------------------------------------------------
# hash grid
# ---------------------------------------------------------------------------------------------
class HashGridSpec:
    """Host-side description of one multi-resolution grid (levels + optional Fourier prefix)."""

    def __init__(self, res: Sequence[int], rows: Sequence[int], n_feat: int, frac_mode: int, n_fourier: int):
        self.res = [int(r) for r in res]
        self.rows = [int(r) for r in rows]
        self.n_feat = int(n_feat)
        self.frac_mode = int(frac_mode)
        self.n_fourier = int(n_fourier)

    @property
    def n_levels(self):
        return len(self.res)

    @property
    def width(self):
        return (3 + 2 * self.n_fourier if self.n_fourier > 0 else 0) + self.n_levels * self.n_feat

    def desc(self, tables: Sequence[torch.Tensor], B: Optional[torch.Tensor]) -> HashGridDesc:
        d = HashGridDesc()
        d.n_levels, d.n_feat, d.frac_mode, d.n_fourier = self.n_levels, self.n_feat, self.frac_mode, self.n_fourier
        for l in range(self.n_levels):
            t = tables[l]
            _f32c(t, "table")
            if not t.is_contiguous() or tuple(t.shape) != (self.rows[l], self.n_feat):
                raise _lib.IdrkError("level %d table must be contiguous [%d, %d]" % (l, self.rows[l], self.n_feat))
            d.res[l] = float(self.res[l])
            d.rows[l] = self.rows[l]
            d.tables[l] = t.data_ptr()
        if self.n_fourier > 0:
            _f32c(B, "fourier B")
            if not B.is_contiguous() or tuple(B.shape) != (3, self.n_fourier):
                raise _lib.IdrkError("fourier B must be contiguous [3, %d]" % self.n_fourier)
            d.fourier_B = B.data_ptr()
        return d


def hash_encode_fwd(spec: HashGridSpec, x: torch.Tensor, tables, B, out: Optional[torch.Tensor] = None,
                    want_idx: bool = False):
    """K1.  x [n, >=3] -> out [n, pad4(width)] (returned tensor is the padded storage)."""
    x = rows2d(x, "x")
    n = x.shape[0]
    ld = pad4(spec.width)
    if out is None:
        out = torch.empty((n, ld), device=x.device, dtype=torch.float32)
    idx = torch.empty((n, spec.n_levels, 8), device=x.device, dtype=torch.int32) if want_idx else None
    if n == 0:
        return (out, idx) if want_idx else out
    d = spec.desc(tables, B)
    check(lib().idrk_hash_encode_fwd(ctypes.byref(d), ptr(x), n, ld_of(x), ptr(out), ld_of(out),
                                     ptr(idx), stream_ptr()), "idrk_hash_encode_fwd")
    return (out, idx) if want_idx else out


def hash_encode_bwd(spec: HashGridSpec, x: torch.Tensor, tables, B, dy: torch.Tensor,
                    grad_tables: Optional[List[torch.Tensor]], want_dx: bool):
    """K2.  Accumulates into grad_tables (list of [T_l, F], may be None) and returns dx [n,3] or None."""
    x = rows2d(x, "x")
    dy = rows2d(dy, "dy")
    n = x.shape[0]
    dx = torch.empty((n, 3), device=x.device, dtype=torch.float32) if want_dx else None
    if n == 0:
        return dx
    d = spec.desc(tables, B)
    if grad_tables is not None:
        arr = (ctypes.c_void_p * spec.n_levels)(*[g.data_ptr() for g in grad_tables])
        garg = ctypes.cast(arr, ctypes.POINTER(ctypes.c_void_p))
    else:
        garg = ctypes.cast(None, ctypes.POINTER(ctypes.c_void_p))
    check(lib().idrk_hash_encode_bwd(ctypes.byref(d), ptr(x), n, ld_of(x), ptr(dy), ld_of(dy),
                                     garg, ptr(dx), stream_ptr()), "idrk_hash_encode_bwd")
    return dx


# ---------------------------------------------------------------------------------------------
# positional encoding
# ---------------------------------------------------------------------------------------------
def posenc_width(d: int, n_bands: int, include_input: bool) -> int:
    return d * ((2 if include_input else 0) + 2 * n_bands)


def posenc_fwd(x: torch.Tensor, bands: Sequence[float], include_input: bool) -> torch.Tensor:
    x = rows2d(x, "x")
    n, d = x.shape
    w = posenc_width(d, len(bands), include_input)
    out = torch.empty((n, pad4(w)), device=x.device, dtype=torch.float32)
    if n == 0:
        return out
    hb = (ctypes.c_float * max(len(bands), 1))(*bands)
    check(lib().idrk_posenc_fwd(ptr(x), n, d, ld_of(x), hb, len(bands), int(include_input), ptr(out), ld_of(out),
                                stream_ptr()), "idrk_posenc_fwd")
    return out


def posenc_bwd(x: torch.Tensor, bands: Sequence[float], include_input: bool, dy: torch.Tensor) -> torch.Tensor:
    x = rows2d(x, "x")
    dy = rows2d(dy, "dy")
    n, d = x.shape
    dx = torch.empty((n, d), device=x.device, dtype=torch.float32)
    if n == 0:
        return dx
    hb = (ctypes.c_float * max(len(bands), 1))(*bands)
    check(lib().idrk_posenc_bwd(ptr(x), n, d, ld_of(x), hb, len(bands), int(include_input), ptr(dy), ld_of(dy),
                                ptr(dx), ld_of(dx), stream_ptr()), "idrk_posenc_bwd")
    return dx
