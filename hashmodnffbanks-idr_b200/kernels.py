"""Thin typed wrappers around the C ABI (include/idrk.h): tensors in, kernels launched on the
current CUDA stream.  No arithmetic happens here and nothing falls back to PyTorch."""
import ctypes
import math
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import HashGridDesc, PROFILE, check, lib, ptr, require_cuda, stream_ptr


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
        self.frac_mode = int(frac_mode)
        # reference / 8-corner modes: integer resolutions; HASH_NGP: the (float) per-level scale base * s^l - 1
        self.res = [float(r) if self.frac_mode == _lib.HASH_NGP else int(r) for r in res]
        self.rows = [int(r) for r in rows]
        self.n_feat = int(n_feat)
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


_SORT_WS: dict = {}


def morton_perm(x: torch.Tensor, lo=None, hi=None, bits: int = 8) -> torch.Tensor:
    """int32 [n] permutation that walks the points of x [n, >=3] in Z-order over the box [lo, hi] - by default the batch's
    own bounding box, measured on the device (csrc/point_sort.cu: own LSD radix sort, ceil(3 * bits / 8) passes; no host
    synchronisation).  Feed it to hash_encode_fwd / hash_encode_bwd as `perm`."""
    x = rows2d(x, "x")
    n = x.shape[0]
    perm = torch.empty(n, device=x.device, dtype=torch.int32)
    if n == 0:
        return perm
    need = ctypes.c_int64(0)
    check(lib().idrk_morton_sort_workspace(n, ctypes.byref(need)), "idrk_morton_sort_workspace")
    ws = _SORT_WS.get(x.device)
    if ws is None or ws.numel() < need.value:
        retire_scratch(ws)                   # a CUDA graph may hold its address (see SCRATCH_GENERATION below)
        ws = _SORT_WS[x.device] = torch.empty(int(need.value * 1.1) + 256, device=x.device, dtype=torch.uint8)
    if (lo is None) != (hi is None):
        raise _lib.IdrkError("morton_perm: give both lo and hi, or neither")
    h_lo = (ctypes.c_float * 3)(*[float(v) for v in lo]) if lo is not None else None
    h_hi = (ctypes.c_float * 3)(*[float(v) for v in hi]) if hi is not None else None
    check(lib().idrk_morton_sort(ptr(x), n, ld_of(x), h_lo, h_hi, int(bits), ptr(perm), ptr(ws), ws.numel(), stream_ptr()),
          "idrk_morton_sort")
    return perm


def _perm_arg(perm, n):
    if perm is None:
        return None
    if perm.dtype != torch.int32 or not perm.is_cuda or not perm.is_contiguous() or perm.numel() != n:
        raise _lib.IdrkError("perm must be a contiguous CUDA int32 tensor with one entry per point")
    return perm


def hash_encode_fwd(spec: HashGridSpec, x: torch.Tensor, tables, B, out: Optional[torch.Tensor] = None,
                    want_idx: bool = False, m_count: Optional[torch.Tensor] = None, rows: Optional[int] = None,
                    perm: Optional[torch.Tensor] = None):
    """K1.  x [n, >=3] -> out [n, pad4(width)] (returned tensor is the padded storage).  `perm` (morton_perm): processing
    order of the points; rows of x / out keep their positions."""
    x = rows2d(x, "x")
    n = x.shape[0] if rows is None else rows
    ld = pad4(spec.width)
    if out is None:
        out = torch.empty((n, ld), device=x.device, dtype=torch.float32)
    idx = torch.empty((n, spec.n_levels, 8), device=x.device, dtype=torch.int32) if want_idx else None
    if n == 0:
        return (out, idx) if want_idx else out
    d = spec.desc(tables, B)
    check(lib().idrk_hash_encode_fwd(ctypes.byref(d), ptr(x), n, ld_of(x), ptr(out), ld_of(out),
                                     ptr(idx), ptr(m_count), ptr(_perm_arg(perm, n)), stream_ptr()), "idrk_hash_encode_fwd")
    return (out, idx) if want_idx else out


def hash_encode_f16pair(spec: HashGridSpec, x: torch.Tensor, tables, B, rows: int, h: torch.Tensor, l: torch.Tensor,
                        ld_out: int, pad_cols: int = 0, m_count: Optional[torch.Tensor] = None, second=None):
    """K1p: the embedding of the first min(rows, *m_count) points written directly as the fp16 pair the SDF pipeline's
    contraction consumes (and, with `second = (h2, l2, ld_out2, pad_cols2, scale2)`, a scaled second copy)."""
    x = rows2d(x, "x")
    if rows == 0:
        return
    h2, l2, ld2, pad2, scale2 = second if second is not None else (None, None, 0, 0, 0.0)
    d = spec.desc(tables, B)
    check(lib().idrk_hash_encode_f16pair(ctypes.byref(d), ptr(x), rows, ld_of(x), ptr(m_count), ptr(h), ptr(l), ld_out,
                                         pad_cols, ptr(h2), ptr(l2), ld2, pad2, float(scale2), stream_ptr()),
          "idrk_hash_encode_f16pair")


HASH_BWD_ORDERED = 1
# set_deterministic_table_grads(True): table gradients of every hash-grid backward are summed in a fixed order (K2d)
DETERMINISTIC_TABLE_GRADS = [False]


def set_deterministic_table_grads(on: bool):
    DETERMINISTIC_TABLE_GRADS[0] = bool(on)


def hash_encode_bwd(spec: HashGridSpec, x: torch.Tensor, tables, B, dy: torch.Tensor,
                    grad_tables: Optional[List[torch.Tensor]], want_dx: bool, ordered: bool = False,
                    perm: Optional[torch.Tensor] = None, deterministic: Optional[bool] = None):
    """K2.  Accumulates into grad_tables (list of [T_l, F], may be None) and returns dx [n,3] or None.  `ordered`: the
    points are spatially ordered (ray samples, utils.sorting.morton_order) - runs sharing a cell are merged in registers."""
    x = rows2d(x, "x")
    dy = rows2d(dy, "dy")
    n = x.shape[0]
    dx = torch.empty((n, 3), device=x.device, dtype=torch.float32) if want_dx else None
    if n == 0:
        return dx
    d = spec.desc(tables, B)
    if grad_tables is not None and spec.n_levels > 0 and (deterministic or (deterministic is None and DETERMINISTIC_TABLE_GRADS[0])):
        arr = (ctypes.c_void_p * spec.n_levels)(*[g.data_ptr() for g in grad_tables])
        need = ctypes.c_int64(0)
        check(lib().idrk_hash_encode_bwd_det_workspace(ctypes.byref(d), n, ctypes.byref(need)), "idrk_hash_encode_bwd_det_workspace")
        ws = _SORT_WS.get(("det", x.device))
        if ws is None or ws.numel() < need.value:
            retire_scratch(ws)
            ws = _SORT_WS[("det", x.device)] = torch.empty(int(need.value * 1.1) + 256, device=x.device, dtype=torch.uint8)
        check(lib().idrk_hash_encode_bwd_det(ctypes.byref(d), ptr(x), n, ld_of(x), ptr(dy), ld_of(dy),
                                             ctypes.cast(arr, ctypes.POINTER(ctypes.c_void_p)), ptr(ws), ws.numel(), stream_ptr()),
              "idrk_hash_encode_bwd_det")
        grad_tables = None
        if not want_dx:
            return None
    if grad_tables is not None:
        arr = (ctypes.c_void_p * spec.n_levels)(*[g.data_ptr() for g in grad_tables])
        garg = ctypes.cast(arr, ctypes.POINTER(ctypes.c_void_p))
    else:
        garg = ctypes.cast(None, ctypes.POINTER(ctypes.c_void_p))
    check(lib().idrk_hash_encode_bwd(ctypes.byref(d), ptr(x), n, ld_of(x), ptr(dy), ld_of(dy),
                                     garg, ptr(dx), HASH_BWD_ORDERED if ordered else 0, ptr(_perm_arg(perm, n)), stream_ptr()),
          "idrk_hash_encode_bwd")
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


def posenc_dx_bwd(g: torch.Tensor, x: torch.Tensor, bands: Sequence[float], include_input: bool, dy: torch.Tensor,
                  want_gdy: bool = True, want_gx: bool = True):
    """Backward of posenc_bwd's map (x, dy) -> dx: (g_dy [n, width] or None, g_x [n, d] or None) for g = dL/d dx."""
    g, x, dy = rows2d(g, "g"), rows2d(x, "x"), rows2d(dy, "dy")
    n, d = x.shape
    width = posenc_width(d, len(bands), include_input)
    g_dy = torch.empty((n, pad4(width)), device=x.device, dtype=torch.float32) if want_gdy else None
    g_x = torch.empty((n, d), device=x.device, dtype=torch.float32) if want_gx else None
    if n == 0 or (g_dy is None and g_x is None):
        return (g_dy[:, :width] if g_dy is not None else None), g_x
    hb = (ctypes.c_float * max(len(bands), 1))(*bands)
    check(lib().idrk_posenc_dx_bwd(ptr(g), ld_of(g), ptr(x), n, d, ld_of(x), hb, len(bands), int(include_input), ptr(dy), ld_of(dy),
                                   ptr(g_dy), pad4(width) if g_dy is not None else 0, ptr(g_x), d, stream_ptr()), "idrk_posenc_dx_bwd")
    return (g_dy[:, :width] if g_dy is not None else None), g_x


# ---------------------------------------------------------------------------------------------
# MLP contraction tiles
# ---------------------------------------------------------------------------------------------
from ._lib import (Epilogue, GEMM_NT, GEMM_NN, GEMM_TN, PREC_FP32, PREC_TF32, PREC_3XTF32,  # noqa: E402
                   EPI_NONE, EPI_SOFTPLUS, EPI_RELU, EPI_MUL_AUX, EPI_SINE, EPI_TANH)

_PRECISION = {"fp32": PREC_FP32, "tf32": PREC_TF32, "3xtf32": PREC_3XTF32}
_default_precision = PREC_3XTF32


def set_precision(name: str):
    """'3xtf32' (default, fp32-accurate tensor-core split), 'tf32' (one pass) or 'fp32' (FFMA tiles)."""
    global _default_precision
    _default_precision = _PRECISION[name]


def get_precision() -> int:
    return _default_precision


def empty_padded(rows: int, cols: int, device) -> torch.Tensor:
    """[rows, cols] view of a [rows, pad4(cols)] buffer (16-byte aligned rows for TMA / float4)."""
    return torch.empty((rows, pad4(cols)), device=device, dtype=torch.float32)[:, :cols]


def operand(t: torch.Tensor, name: str = "operand") -> torch.Tensor:
    """2-D fp32 CUDA tensor usable by the tensor-core path: unit column stride, ld % 4 == 0, 16B base."""
    _f32c(t, name)
    if t.dim() != 2:
        raise _lib.IdrkError("%s must be 2-D" % name)
    ok = t.stride(1) == 1 and (t.shape[0] <= 1 or (t.stride(0) % 4 == 0 and t.stride(0) >= t.shape[1])) \
        and t.data_ptr() % 16 == 0
    if ok:
        return t
    buf = empty_padded(t.shape[0], t.shape[1], t.device)
    buf.copy_(t)
    return buf


def op_ld(t: torch.Tensor) -> int:
    return t.stride(0) if t.shape[0] > 1 else pad4(t.shape[1])


def split_tf32(x: torch.Tensor, m_count: Optional[torch.Tensor] = None):
    """(hi, lo) with hi = tf32(x), lo = tf32(x - hi); both padded like `operand`."""
    x = rows2d(x, "x")
    r, c = x.shape
    hi = empty_padded(r, c, x.device)
    lo = empty_padded(r, c, x.device)
    if r:
        check(lib().idrk_split_tf32(ptr(x), r, c, ld_of(x), 1.0, ptr(hi), ptr(lo), pad4(c), pad4(c) - c, ptr(m_count),
                                    stream_ptr()), "idrk_split_tf32")
    return hi, lo


def split_into(x: torch.Tensor, rows: int, cols: int, scale: float, hi: torch.Tensor, lo: Optional[torch.Tensor],
               ld_out: int, pad_cols: int = 0, m_count: Optional[torch.Tensor] = None):
    """Raw form: writes scale*x (split when `lo` is given) into caller-owned buffers (column offsets via views)."""
    check(lib().idrk_split_tf32(ptr(x), rows, cols, ld_of(x), float(scale), ptr(hi), ptr(lo), ld_out, pad_cols,
                                ptr(m_count), stream_ptr()), "idrk_split_tf32")


def gemm(layout: int, A: torch.Tensor, B: torch.Tensor, M: int, N: int, Kc: int, *, precision: Optional[int] = None,
         A_lo=None, B_lo=None, C=None, C_hi=None, C_lo=None, S=None, bias=None, aux=None, mode=EPI_NONE,
         act=0.0, scale=1.0, accumulate=False, m_count=None, split_k=1):
    """Raw launcher.  All operands must already satisfy `operand()`; output buffers are caller-owned."""
    prec = _default_precision if precision is None else precision
    e = Epilogue()
    outs = [t for t in (C, C_hi, C_lo) if t is not None]
    if not outs:
        raise _lib.IdrkError("gemm needs an output")
    e.C, e.C_hi, e.C_lo, e.S = (t.data_ptr() if t is not None else None for t in (C, C_hi, C_lo, S))
    e.bias = bias.data_ptr() if bias is not None else None
    e.aux = aux.data_ptr() if aux is not None else None
    e.ldc = op_ld(outs[0])
    e.lds = op_ld(S) if S is not None else 0
    e.ldaux = op_ld(aux) if aux is not None else 0
    e.mode, e.act_param, e.scale, e.accumulate = mode, float(act), float(scale), int(bool(accumulate))
    if PROFILE.enabled:
        PROFILE.pending_tag = "[%s M=%d N=%d K=%d mode=%d%s%s]" % ("NT NN TN".split()[layout], M, N, Kc, mode,
                                                                  " cnt" if m_count is not None else "",
                                                                  " split%d" % split_k if split_k > 1 else "")
        if layout == GEMM_NT and M >= 16384 and N >= 256 and split_k == 1:
            PROFILE.pending_tag = "_2cta" + PROFILE.pending_tag
        if m_count is not None:
            cnt = m_count.clone()
            PROFILE.pending_flops = lambda cnt=cnt, M=M, N=N, Kc=Kc: 2.0 * min(M, int(cnt.item())) * N * Kc
        else:
            PROFILE.pending_flops = 2.0 * M * N * Kc
    check(lib().idrk_gemm(layout, prec, M, N, Kc, ptr(A), ptr(A_lo), op_ld(A), ptr(B), ptr(B_lo), op_ld(B),
                          ctypes.byref(e), ptr(m_count), split_k, stream_ptr()), "idrk_gemm")


def weight_norm_fwd(g: Optional[torch.Tensor], v: torch.Tensor, want_split: bool, want_t: bool, out=None):
    """Returns dict with W (+ W_hi/W_lo, Wt (+ Wt_hi/Wt_lo)), all padded operands.  `out` reuses the buffers of a
    previous call (in-place refresh: pointers stay valid for CUDA graphs)."""
    v = rows2d(v, "weight_v")
    N, Kd = v.shape
    dev = v.device
    if out is None:
        out = {"W": empty_padded(N, Kd, dev)}
        if want_split:
            out["W_hi"], out["W_lo"] = empty_padded(N, Kd, dev), empty_padded(N, Kd, dev)
        if want_t:
            out["Wt"] = empty_padded(Kd, N, dev)
            if want_split:
                out["Wt_hi"], out["Wt_lo"] = empty_padded(Kd, N, dev), empty_padded(Kd, N, dev)
    gg = g.reshape(-1).contiguous() if g is not None else None
    check(lib().idrk_weight_norm_fwd(ptr(gg), ptr(v), N, Kd, ld_of(v), ptr(out["W"]), ptr(out.get("W_hi")), ptr(out.get("W_lo")),
                                     pad4(Kd), ptr(out.get("Wt")), ptr(out.get("Wt_hi")), ptr(out.get("Wt_lo")), pad4(N),
                                     stream_ptr()), "idrk_weight_norm_fwd")
    return out


def weight_norm_bwd(g: torch.Tensor, v: torch.Tensor, dW: torch.Tensor, into=None):
    """(dg, dv) of W = g v / ||v||.  `into=(dg_buf, dv_buf)` ACCUMULATES into caller-owned contiguous buffers (the
    parameters' .grad views of the flat bucket) instead of allocating."""
    v = rows2d(v, "weight_v")
    dW = rows2d(dW, "dW")
    N, Kd = v.shape
    if into is None:
        dg = torch.empty((N, 1), device=v.device, dtype=torch.float32)
        dv = torch.empty((N, Kd), device=v.device, dtype=torch.float32)
        acc = 0
    else:
        dg, dv = into
        if not (dg.is_contiguous() and dv.is_contiguous() and dg.numel() == N and dv.shape == (N, Kd)):
            raise _lib.IdrkError("weight_norm_bwd: accumulation buffers must be contiguous [N,1] / [N,K]")
        acc = 1
    check(lib().idrk_weight_norm_bwd(ptr(g.reshape(-1).contiguous()), ptr(v), ptr(dW), N, Kd, ld_of(v), ld_of(dW),
                                     ptr(dg), ptr(dv), Kd, acc, stream_ptr()), "idrk_weight_norm_bwd")
    return dg, dv


# ---------------------------------------------------------------------------------------------
# scratch buffers whose addresses CUDA graphs hold
# ---------------------------------------------------------------------------------------------
# The ray tracer's and the trainer's CUDA graphs are captured against eagerly allocated, shared scratch (SdfPipeline
# buffers, the zero arena) and keep raw pointers.  A buffer that has to grow after a capture is therefore never
# freed: the old tensor is parked in `_RETIRED` (the graphs that point at it stay valid) and `SCRATCH_GENERATION`
# moves, which is part of every graph key - holders drop their graph and recapture against the new buffers.
SCRATCH_GENERATION = [0]
GRAPHS_CAPTURED = [0]
_RETIRED: list = []


def note_graph_captured():
    GRAPHS_CAPTURED[0] += 1


def retire_scratch(t):
    """Called instead of dropping a scratch tensor that is being replaced by a larger one."""
    if t is None:
        return
    if GRAPHS_CAPTURED[0] > 0:
        _RETIRED.append(t)
        SCRATCH_GENERATION[0] += 1


class ZeroPool:
    """Arena of pre-zeroed fp32 scratch for one differentiable step.

    Split-K outputs, column sums and hash-table gradients all accumulate with atomics into zero-initialised buffers;
    a step needs ~150 of them, and one `torch.zeros` each is one fill kernel each.  Between `begin()` and `end()`
    `take()` hands out slices of ONE buffer zeroed by ONE fill; outside (or when the arena is too small - it grows
    for the next step) it falls back to `torch.zeros`.  Slices are only valid until the next `begin()`."""

    def __init__(self):
        self.buf = None
        self.off = 0
        self.need = 0
        self.active = False

    def begin(self, device):
        want = max(self.need, 1 << 20)
        if self.buf is None or self.buf.device != device or self.buf.numel() < want:
            if torch.cuda.is_current_stream_capturing():
                raise _lib.IdrkError("ZeroPool would have to grow inside a CUDA-graph capture (warm the step up eagerly first)")
            retire_scratch(self.buf)
            self.buf = torch.empty(int(want * 1.25), device=device, dtype=torch.float32)
        self.buf.zero_()
        self.off = 0
        self.need = 0
        self.active = True

    def end(self):
        self.active = False

    def take(self, numel: int, device) -> torch.Tensor:
        n = (numel + 63) // 64 * 64                      # 256-byte aligned slices
        self.need += n
        if not self.active or self.buf is None or self.buf.device != device or self.off + n > self.buf.numel():
            return torch.zeros(numel, device=device, dtype=torch.float32)
        t = self.buf[self.off:self.off + numel]
        self.off += n
        return t


ZERO_POOL = ZeroPool()

# Trainer mode (DataParallelTrainer sets it around .backward()): leaf gradients whose producer is one of our kernels
# are accumulated by that kernel straight into the parameter's .grad (a view of the flat bucket) and the autograd
# Function returns None for them - same result as AccumulateGrad (.grad += g) without the temporary and the add.
DIRECT_GRADS = [False]


def direct_grad_target(p):
    if DIRECT_GRADS[0] and not torch.is_grad_enabled() and p is not None and p.is_leaf and p.grad is not None \
            and p.grad.is_contiguous():
        return p.grad
    return None


def colsum(x: torch.Tensor, into: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[c] (+)= sum_r x[r, c]; `into` accumulates into a caller-owned zero-initialised / gradient buffer."""
    x = rows2d(x, "x")
    out = ZERO_POOL.take(x.shape[1], x.device) if into is None else into
    if x.shape[0]:
        check(lib().idrk_colsum(ptr(x), x.shape[0], x.shape[1], ld_of(x), ptr(out), stream_ptr()), "idrk_colsum")
    return out


def sdf_head(h: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, beta: float, out: torch.Tensor, rows: int,
             m_count=None):
    check(lib().idrk_sdf_head(ptr(h), rows, w.numel(), op_ld(h), ptr(w), ptr(bias), float(beta), ptr(out), ptr(m_count),
                              stream_ptr()), "idrk_sdf_head")


def sdf_squash_rows(x: torch.Tensor, beta: float, want_grads: bool):
    """out = x with column 0 squashed (padded operand), d = d out_0/ds [n,1], d2 = second derivative [n,1]."""
    x = rows2d(x, "x")
    n, cols = x.shape
    out = empty_padded(n, cols, x.device)
    d = torch.empty((n, 1), device=x.device, dtype=torch.float32) if want_grads else None
    d2 = torch.empty((n, 1), device=x.device, dtype=torch.float32) if want_grads else None
    if n:
        check(lib().idrk_sdf_squash_rows(ptr(x), n, cols, ld_of(x), float(beta), ptr(out), pad4(cols), ptr(d), ptr(d2), stream_ptr()),
              "idrk_sdf_squash_rows")
    return out, d, d2


def sdf_squash(s: torch.Tensor, beta: float, want_grad: bool):
    s = s.contiguous()
    out = torch.empty_like(s)
    d = torch.empty_like(s) if want_grad else None
    if s.numel():
        check(lib().idrk_sdf_squash(ptr(s), s.numel(), float(beta), ptr(out), ptr(d), stream_ptr()), "idrk_sdf_squash")
    return out, d


# ---------------------------------------------------------------------------------------------
# O(rays) ends of the step: camera rays + sphere intersection, IDR loss (csrc/render_glue.cu)
# ---------------------------------------------------------------------------------------------
def camera_rays(uv: torch.Tensor, pose: torch.Tensor, intrinsics: torch.Tensor, radius: Optional[float] = None):
    """uv [B,N,2], pose [B,4,4], intrinsics [B,4,4] -> (ray_dirs [B,N,3], cam_loc [B,3]) and, with `radius`, also
    (t_sph [B,N,2], hit [B,N] bool) of the bounding sphere - one launch, no boolean indexing (no host sync)."""
    for t, name in ((uv, "uv"), (pose, "pose"), (intrinsics, "intrinsics")):
        _f32c(t, name)
    B, N = uv.shape[0], uv.shape[1]
    if tuple(pose.shape) != (B, 4, 4) or tuple(intrinsics.shape) != (B, 4, 4) or uv.shape[2] != 2:
        raise _lib.IdrkError("camera_rays: uv [B,N,2], pose [B,4,4], intrinsics [B,4,4] expected")
    uv, pose, intrinsics = uv.contiguous(), pose.contiguous(), intrinsics.contiguous()
    dev = uv.device
    dirs = torch.empty((B, N, 3), device=dev, dtype=torch.float32)
    cam = torch.empty((B, 3), device=dev, dtype=torch.float32)
    t_sph = hit = None
    if radius is not None:
        t_sph = torch.empty((B, N, 2), device=dev, dtype=torch.float32)
        hit = torch.empty((B, N), device=dev, dtype=torch.bool)
    check(lib().idrk_camera_rays(ptr(uv), ptr(pose), ptr(intrinsics), B, N, float(radius if radius is not None else 1.0),
                                 ptr(dirs), ptr(cam), ptr(t_sph), ptr(hit), stream_ptr()), "idrk_camera_rays")
    return (dirs, cam) if radius is None else (dirs, cam, t_sph, hit)


def idr_loss(rgb_values, rgb_gt, net_mask, obj_mask, sdf_output, grad_theta, eikonal_weight, mask_weight, alpha,
             want_grads: bool):
    """One launch: out[4] = (loss, rgb_loss, eikonal_loss, mask_loss) and (want_grads) d loss / d (rgb_values, sdf_output,
    grad_theta)."""
    rgb_values = rows2d(rgb_values, "rgb_values")
    sdf_output = rows2d(sdf_output.reshape(-1, 1), "sdf_output")
    n = rgb_values.shape[0]
    rgb_gt = _f32c(rgb_gt, "rgb_gt").reshape(-1, 3).contiguous()
    nm, om = net_mask.reshape(-1).contiguous(), obj_mask.reshape(-1).contiguous()
    if nm.dtype != torch.bool or om.dtype != torch.bool or nm.numel() != n or om.numel() != n or rgb_gt.shape[0] != n:
        raise _lib.IdrkError("idr_loss: masks must be bool [N] and rgb_gt [N,3]")
    m = 0
    if grad_theta is not None and grad_theta.shape[0] > 0:
        grad_theta = rows2d(grad_theta, "grad_theta")
        m = grad_theta.shape[0]
    dev = rgb_values.device
    out = torch.empty(4, device=dev, dtype=torch.float32)
    d_rgb = torch.empty((n, 3), device=dev, dtype=torch.float32) if want_grads else None
    d_sdf = torch.empty((n, 1), device=dev, dtype=torch.float32) if want_grads else None
    d_g = torch.empty((m, 3), device=dev, dtype=torch.float32) if (want_grads and m) else None
    check(lib().idrk_idr_loss(ptr(rgb_values), ld_of(rgb_values), ptr(rgb_gt), ptr(nm), ptr(om), ptr(sdf_output),
                              ld_of(sdf_output), n, ptr(grad_theta) if m else None, ld_of(grad_theta) if m else 0, m,
                              float(eikonal_weight), float(mask_weight), float(alpha), ptr(out), ptr(d_rgb), ptr(d_sdf),
                              ptr(d_g), stream_ptr()), "idrk_idr_loss")
    return out, d_rgb, d_sdf, d_g


def fourier_dx_fwd(x: torch.Tensor, dy: torch.Tensor, B: torch.Tensor) -> torch.Tensor:
    x, dy = rows2d(x, "x"), rows2d(dy, "dy")
    dx = torch.empty((x.shape[0], 3), device=x.device, dtype=torch.float32)
    check(lib().idrk_fourier_dx_fwd(ptr(x), ld_of(x), ptr(dy), ld_of(dy), ptr(B), B.shape[1], x.shape[0], ptr(dx), stream_ptr()),
          "idrk_fourier_dx_fwd")
    return dx


def fourier_dx_bwd(g: torch.Tensor, x: torch.Tensor, dy: torch.Tensor, B: torch.Tensor, width: int, want_gx: bool):
    g, x, dy = rows2d(g, "g"), rows2d(x, "x"), rows2d(dy, "dy")
    n = x.shape[0]
    g_dy = torch.empty((n, pad4(width)), device=x.device, dtype=torch.float32)
    g_x = torch.empty((n, 3), device=x.device, dtype=torch.float32) if want_gx else None
    check(lib().idrk_fourier_dx_bwd(ptr(g), ld_of(g), ptr(x), ld_of(x), ptr(dy), ld_of(dy), ptr(B), B.shape[1], n, ptr(g_dy),
                                    pad4(width), width, ptr(g_x), stream_ptr()), "idrk_fourier_dx_bwd")
    return g_dy[:, :width], g_x


def scale3(scale: torch.Tensor, a, b, c):
    """(scale * a, scale * b, scale * c) in one launch; scale is a 0-dim / 1-element device tensor, c may be None."""
    scale = _f32c(scale.reshape(1), "scale")
    ya, yb = torch.empty_like(a), torch.empty_like(b)
    yc = torch.empty_like(c) if c is not None else None
    check(lib().idrk_scale3(ptr(scale), ptr(a), ptr(ya), a.numel(), ptr(b), ptr(yb), b.numel(), ptr(c), ptr(yc),
                            c.numel() if c is not None else 0, stream_ptr()), "idrk_scale3")
    return ya, yb, yc


# ---------------------------------------------------------------------------------------------
# optimiser on the flat bucket
# ---------------------------------------------------------------------------------------------
def sumsq(g: torch.Tensor, out: torch.Tensor):
    check(lib().idrk_sumsq(ptr(g), g.numel(), ptr(out), stream_ptr()), "idrk_sumsq")


def sumsq_det(g: torch.Tensor, out: torch.Tensor, partials: torch.Tensor):
    """out[0] = sum(g^2), summed in an order fixed by (g.numel(), partials.numel()): identical on every replica."""
    check(lib().idrk_sumsq_det(ptr(g), g.numel(), ptr(out), ptr(partials), partials.numel(), stream_ptr()), "idrk_sumsq_det")


def clip_adam(p, g, m, v, lr, beta1, beta2, eps, step, max_norm, sumsq_buf, grad_scale):
    check(lib().idrk_clip_adam(ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), float(lr), float(beta1), float(beta2), float(eps),
                               int(step), float(max_norm), ptr(sumsq_buf), float(grad_scale), stream_ptr()), "idrk_clip_adam")


def act_bwd(dH, dS, S, H, mode: int, act: float, scale: float, want_split: bool):
    """dZ = dH * S * scale + dS * act''(Z) (+ its hi/lo operand pair)."""
    S = rows2d(S, "S")
    rows, cols = S.shape
    dev = S.device
    dZ = empty_padded(rows, cols, dev)
    hi = empty_padded(rows, cols, dev) if want_split else None
    lo = empty_padded(rows, cols, dev) if want_split else None
    if rows == 0:
        return dZ, hi, lo
    dH = rows2d(dH, "dH") if dH is not None else None
    dS = rows2d(dS, "dS") if dS is not None else None
    H = rows2d(H, "H") if H is not None else None
    check(lib().idrk_act_bwd(ptr(dH), ld_of(dH) if dH is not None else 0, ptr(dS), ld_of(dS) if dS is not None else 0,
                             ptr(S), ld_of(S), ptr(H), ld_of(H) if H is not None else 0, rows, cols, mode, float(act),
                             float(scale), ptr(dZ), ptr(hi), ptr(lo), pad4(cols), stream_ptr()), "idrk_act_bwd")
    return dZ, hi, lo


# ---------------------------------------------------------------------------------------------
# fp16-pair contraction (no-grad SDF path)
# ---------------------------------------------------------------------------------------------
from ._lib import EpilogueH  # noqa: E402

_inference_fp16x2 = True


def set_inference_precision(name: str):
    """Operand format of the no-grad SDF pipeline (ray tracer / eval queries): "fp16x2" = fp16 pairs
    x ~= h + l * 2^-11 (default, 3xTF32-class accuracy at half the bytes and MMA time), "default" = follow
    set_precision()."""
    global _inference_fp16x2
    if name not in ("fp16x2", "default"):
        raise ValueError(name)
    _inference_fp16x2 = name == "fp16x2"


def inference_fp16x2() -> bool:
    return _inference_fp16x2 and _default_precision == PREC_3XTF32


def pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def empty_half(rows: int, cols: int, device) -> torch.Tensor:
    return torch.empty((rows, pad8(cols)), device=device, dtype=torch.float16)[:, :cols]


def _nffb_desc(ffb):
    """idrk_nffb_t of a FourierFilterBanks module (model/embeddings/nffb3d.py) + the tensors it points into."""
    W = ffb.feature_Vector_size
    grid = ffb.grid_enc
    d = _lib.NffbDesc()
    d.grid = grid.spec().desc(tuple(t.detach() for t in grid.tables()), grid.freq_encoding.B)
    bands = ffb.ff_enc[0].freq_bands
    for i, b in enumerate(bands):
        d.bands[i] = float(b)
    d.n_bands, d.include_input = len(bands), 1 if ffb.include_input else 0
    d.n_lin, d.width, d.chunk = ffb.n_nffb_layers - 1, W, 2 * ffb.max_points_per_level
    d.style, d.n_levels_div = (1 if ffb.modulationApplied else 0), ffb.grid_levels
    d.bound, d.w0 = float(ffb.bound), float(ffb.sin_w0)
    keep = []
    for j in range(ffb.n_nffb_layers - 1):
        lin = getattr(ffb, "ff_lin%d" % j)
        w, b = lin.weight.detach().contiguous(), lin.bias.detach().contiguous()
        keep += [w, b]
        d.lin_w[j], d.lin_b[j] = w.data_ptr(), b.data_ptr()
    ow, ob = ffb.out_layer.weight.detach().contiguous(), ffb.out_layer.bias.detach().contiguous()
    keep += [ow, ob]
    d.out_w, d.out_b = ow.data_ptr(), ob.data_ptr()
    if ffb.modulationApplied:
        st = ffb.StyleAttentionBlock
        sw, sb = st.linear_transform.weight.detach().contiguous(), st.linear_transform.bias.detach().contiguous()
        keep += [sw, sb]
        d.style_w, d.style_b, d.eps = sw.data_ptr(), sb.data_ptr(), float(st.eps)
    return d, keep


def nffb_encode_fwd(ffb, x: torch.Tensor, out: Optional[torch.Tensor] = None, m_count: Optional[torch.Tensor] = None,
                    rows: Optional[int] = None) -> torch.Tensor:
    """K7: the whole FourierFilterBanks forward (no autograd) in one launch.  `ffb` is the module
    (model/embeddings/nffb3d.py); returns the padded [n, pad4(3 + width)] buffer."""
    x = rows2d(x, "x")
    n = x.shape[0] if rows is None else rows
    ld = pad4(3 + ffb.feature_Vector_size)
    if out is None:
        out = torch.empty((n, ld), device=x.device, dtype=torch.float32)
    if n == 0:
        return out
    d, keep = _nffb_desc(ffb)
    check(lib().idrk_nffb_encode_fwd(ctypes.byref(d), ptr(x), n, ld_of(x), ptr(out), ld_of(out), ptr(m_count), stream_ptr()),
          "idrk_nffb_encode_fwd")
    return out


def nffb_encode_f16pair(ffb, x: torch.Tensor, rows: int, h: torch.Tensor, l: torch.Tensor, ld_out: int, pad_cols: int = 0,
                        m_count: Optional[torch.Tensor] = None, second=None):
    """K7: the filter-bank embedding of the first min(rows, *m_count) points written directly as the fp16 pair the SDF
    pipeline's contraction consumes (and, with `second = (h2, l2, ld_out2, pad_cols2, scale2)`, a scaled second copy)."""
    x = rows2d(x, "x")
    if rows == 0:
        return
    h2, l2, ld2, pad2, scale2 = second if second is not None else (None, None, 0, 0, 0.0)
    d, keep = _nffb_desc(ffb)
    check(lib().idrk_nffb_encode_f16pair(ctypes.byref(d), ptr(x), rows, ld_of(x), ptr(m_count), ptr(h), ptr(l), ld_out, pad_cols,
                                         ptr(h2), ptr(l2), ld2, pad2, float(scale2), stream_ptr()), "idrk_nffb_encode_f16pair")


def nffb_pair_supported(ffb) -> bool:
    return nffb_fused_supported(ffb)


def nffb_fused_supported(ffb) -> bool:
    W = ffb.feature_Vector_size
    return W <= 64 and (ffb.n_nffb_layers - 2) * 2 * ffb.max_points_per_level <= 32 and ffb.n_nffb_layers - 1 <= 16


def split_f16_into(x: torch.Tensor, rows: int, cols: int, scale: float, h: torch.Tensor, l: torch.Tensor, ld_out: int,
                   pad_cols: int = 0, m_count: Optional[torch.Tensor] = None, second=None):
    """`second = (h2, l2, ld_out2, pad_cols2, scale2)`: a second fp16 pair of scale2 * x written in the same pass."""
    h2, l2, ld2, pad2, scale2 = second if second is not None else (None, None, 0, 0, 0.0)
    check(lib().idrk_split_f16(ptr(x), rows, cols, ld_of(x), float(scale), ptr(h), ptr(l), ld_out, pad_cols,
                               ptr(h2), ptr(l2), ld2, pad2, float(scale2), ptr(m_count), stream_ptr()), "idrk_split_f16")


def split_f16(x: torch.Tensor, m_count: Optional[torch.Tensor] = None):
    x = rows2d(x, "x")
    r, c = x.shape
    h, l = empty_half(r, c, x.device), empty_half(r, c, x.device)
    if r:
        split_f16_into(x, r, c, 1.0, h, l, pad8(c), pad8(c) - c, m_count)
    return h, l


def half_ld(t: torch.Tensor) -> int:
    return t.stride(0) if t.shape[0] > 1 else pad8(t.shape[1])


def gemm_f16s(A_h, A_l, B_h, B_l, M: int, N: int, Kc: int, *, C=None, C_h=None, C_l=None, bias=None, mode=EPI_NONE,
              act=0.0, scale=1.0, m_count=None, dot_w=None, dot_out=None):
    """fp16-pair contraction.  With `dot_w` [N] the activated tile is not stored: `dot_out` [M, >= ceil(N/32)] receives one
    fp32 partial of sum_c act(z[r, c]) * dot_w[c] per 32-column group (the fused SDF head)."""
    e = EpilogueH()
    if dot_w is not None:
        if dot_out is None or dot_out.dtype != torch.float32 or dot_out.stride(-1) != 1 or dot_out.shape[0] < M \
                or dot_w.dtype != torch.float32 or dot_w.numel() < N or not dot_w.is_contiguous():
            raise _lib.IdrkError("gemm_f16s: bad dot_w / dot_out")
        e.dot_w, e.dot_out, e.ld_dot = dot_w.data_ptr(), dot_out.data_ptr(), dot_out.stride(0)
    e.C = C.data_ptr() if C is not None else None
    e.C_h = C_h.data_ptr() if C_h is not None else None
    e.C_l = C_l.data_ptr() if C_l is not None else None
    e.bias = bias.data_ptr() if bias is not None else None
    e.ldc = op_ld(C) if C is not None else 0
    e.ldh = half_ld(C_h) if C_h is not None else 0
    e.mode, e.act_param, e.scale = mode, float(act), float(scale)
    if PROFILE.enabled:
        # "_big": the 100-sample sweeps' chunks (32 K rows); the sphere-tracing queries have capacities of 2 N .. 6 N rows
        PROFILE.pending_tag = ("_big" if M >= 16384 else "") + "[f16s M=%d N=%d K=%d mode=%d%s]" % (
            M, N, Kc, mode, " cnt" if m_count is not None else "")
        if m_count is not None:
            cnt = m_count.clone()
            PROFILE.pending_flops = lambda cnt=cnt, M=M, N=N, Kc=Kc: 2.0 * min(M, int(cnt.item())) * N * Kc
        else:
            PROFILE.pending_flops = 2.0 * M * N * Kc
    check(lib().idrk_gemm_f16s(M, N, Kc, ptr(A_h), ptr(A_l), half_ld(A_h), ptr(B_h), ptr(B_l), half_ld(B_h),
                               ctypes.byref(e), ptr(m_count), stream_ptr()), "idrk_gemm_f16s")


# ---------------------------------------------------------------------------------------------
# 16-bit-pair contraction of the DIFFERENTIABLE path (csrc/gemm_p16.cu)
# ---------------------------------------------------------------------------------------------
from ._lib import EpilogueP, P16_FP16, P16_BF16  # noqa: E402

_training_p16 = True

def set_training_operands(name: str):
    """Operand format of the autograd path's contractions under set_precision("3xtf32"): "p16" = pairs of 16-bit floats
    x ~= h + l * 2^-11 (default: bf16 pairs - full fp32 range, ~17 significant bits; 4 bytes per element, kind::f16
    tensor rate) or "tf32x3" = tf32 hi / lo pairs (8 bytes per element, kind::tf32)."""
    global _training_p16
    if name not in ("p16", "tf32x3"):
        raise ValueError(name)
    _training_p16 = name == "p16"


def training_p16() -> bool:
    return _training_p16 and _default_precision == PREC_3XTF32


def empty_pair16(rows: int, cols: int, device, fmt: int):
    dt = torch.float16 if fmt == P16_FP16 else torch.bfloat16
    buf = torch.empty((2, max(rows, 1), pad8(cols)), device=device, dtype=dt)
    return buf[0, :rows, :cols], buf[1, :rows, :cols]


def split_p16(x: torch.Tensor, fmt: int, m_count: Optional[torch.Tensor] = None):
    """(h, l, fmt) pair of a 2-D fp32 tensor."""
    x = rows2d(x, "x")
    r, c = x.shape
    h, l = empty_pair16(r, c, x.device, fmt)
    if r:
        check(lib().idrk_split_p16(ptr(x), r, c, ld_of(x), 1.0, ptr(h), ptr(l), pad8(c), pad8(c) - c, fmt, ptr(m_count),
                                   stream_ptr()), "idrk_split_p16")
    return h, l, fmt


def gemm_p16(layout: int, A, B, M: int, N: int, Kc: int, *, C=None, C_pair=None, S=None, bias=None, aux=None,
             mode=EPI_NONE, act=0.0, scale=1.0, accumulate=False, m_count=None, split_k=1):
    """Raw launcher.  A, B: (h, l, fmt) pairs (rows padded to 8 halves); C_pair: (C_h, C_l, fmt) output pair."""
    e = EpilogueP()
    if C is None and C_pair is None:
        raise _lib.IdrkError("gemm_p16 needs an output")
    e.C = C.data_ptr() if C is not None else None
    e.S = S.data_ptr() if S is not None else None
    if C_pair is not None:
        e.C_h, e.C_l, e.c_fmt = C_pair[0].data_ptr(), C_pair[1].data_ptr(), int(C_pair[2])
        e.ldh = half_ld(C_pair[0])
    e.bias = bias.data_ptr() if bias is not None else None
    e.aux = aux.data_ptr() if aux is not None else None
    e.ldc = op_ld(C) if C is not None else 0
    e.lds = op_ld(S) if S is not None else 0
    e.ldaux = op_ld(aux) if aux is not None else 0
    e.mode, e.act_param, e.scale, e.accumulate = mode, float(act), float(scale), int(bool(accumulate))
    if PROFILE.enabled:
        PROFILE.pending_tag = "[p16 %s M=%d N=%d K=%d mode=%d fmt=%d%d%s]" % ("NT NN TN".split()[layout], M, N, Kc, mode, A[2], B[2],
                                                                           " split%d" % split_k if split_k > 1 else "")
        PROFILE.pending_flops = 2.0 * M * N * Kc
    check(lib().idrk_gemm_p16(layout, M, N, Kc, ptr(A[0]), ptr(A[1]), int(A[2]), half_ld(A[0]), ptr(B[0]), ptr(B[1]), int(B[2]),
                              half_ld(B[0]), ctypes.byref(e), ptr(m_count), split_k, stream_ptr()), "idrk_gemm_p16")


def weight_norm_fwd_p16(g: Optional[torch.Tensor], v: torch.Tensor, fmt: int):
    """W = g v / ||v|| as fp32 and as a 16-bit pair (one launch)."""
    v = rows2d(v, "weight_v")
    N, Kd = v.shape
    W = empty_padded(N, Kd, v.device)
    h, l = empty_pair16(N, Kd, v.device, fmt)
    gg = g.reshape(-1).contiguous() if g is not None else None
    check(lib().idrk_weight_norm_fwd_p16(ptr(gg), ptr(v), N, Kd, ld_of(v), ptr(W), pad4(Kd), ptr(h), ptr(l), pad8(Kd), fmt,
                                         stream_ptr()), "idrk_weight_norm_fwd_p16")
    return W, (h, l, fmt)


def act_bwd_p16(dH, dS, S, H, mode: int, act: float, scale: float, fmt: int):
    """dZ = dH * S * scale + dS * act''(Z) as fp32 and as a 16-bit pair."""
    S = rows2d(S, "S")
    rows, cols = S.shape
    dev = S.device
    dZ = empty_padded(rows, cols, dev)
    h, l = empty_pair16(rows, cols, dev, fmt)
    if rows == 0:
        return dZ, (h, l, fmt)
    dH = rows2d(dH, "dH") if dH is not None else None
    dS = rows2d(dS, "dS") if dS is not None else None
    H = rows2d(H, "H") if H is not None else None
    check(lib().idrk_act_bwd_p16(ptr(dH), ld_of(dH) if dH is not None else 0, ptr(dS), ld_of(dS) if dS is not None else 0,
                                 ptr(S), ld_of(S), ptr(H), ld_of(H) if H is not None else 0, rows, cols, mode, float(act),
                                 float(scale), ptr(dZ), pad4(cols), ptr(h), ptr(l), pad8(cols), fmt, stream_ptr()),
          "idrk_act_bwd_p16")
    return dZ, (h, l, fmt)
