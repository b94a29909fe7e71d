"""torch.autograd.Function wrappers: the only place where kernels meet autograd.

Every Function's forward/backward launches libidrk kernels.  Where the reference needs SECOND-order
derivatives (ImplicitNetwork.gradient uses create_graph=True, implicit_differentiable_renderer.py:116-128)
the backward pass is expressed through other differentiable Functions / device tensor ops whenever
grad mode is on, so double backward works without hand-written second-order kernels.
"""
import math
import os
from typing import List, Optional, Sequence

import torch

from . import kernels as K
from ._lib import IdrkError

TWO_PI = 2.0 * math.pi


# ---------------------------------------------------------------------------------------------
# hash grid (+ Fourier prefix)
# ---------------------------------------------------------------------------------------------
class _FourierDx(torch.autograd.Function):
    """d/dx of the Fourier prefix as ONE kernel that stays differentiable in dy and x (one more kernel): the recorded
    backward pass of ImplicitNetwork.gradient needs it, and ~40 tensor ops (two K = 3 matmuls among them) did it before."""

    @staticmethod
    def forward(ctx, x, dy, B, width):
        ctx.width = width
        ctx.save_for_backward(x, dy, B)
        return K.fourier_dx_fwd(x.detach(), dy.detach(), B)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        x, dy, B = ctx.saved_tensors
        g_dy, g_x = K.fourier_dx_bwd(g, x, dy, B, ctx.width, ctx.needs_input_grad[0])
        if g_x is not None and x.shape[1] != 3:
            pad = torch.zeros_like(x)
            pad[:, :3] = g_x
            g_x = pad
        return g_x, (g_dy if ctx.needs_input_grad[1] else None), None, None


# 8-corner modes on big unordered batches: walk the points in Z-order (one Morton radix sort per forward, shared with the
# backward).  Measured on 2^24 uniform points, L = 16, F = 2, SORT INCLUDED (profiles/r02_hash_encode_sweep_sort_v2.txt,
# fractions of the HBM roofline): forward + table-gradient pair 0.49 -> 0.70 at T = 2^19 and 0.40 -> 0.53 at T = 2^22; the
# forward alone is level while the tables sit in L2 (0.67 -> 0.67) and wins once they do not (T = 2^22: 0.35-0.47 -> 0.60).
# At T = 2^24 (1.2 GB) nothing is gained.  The reference-mode passes (one gather per level) are bound by their row traffic
# and lose from a permuted walk: never sorted.  Hence: sort when a backward will share the permutation, or when the tables
# exceed L2; never beyond 512 MB of tables or below 2^18 points.
AUTO_SORT = {"enabled": os.environ.get("IDRK_HASH_AUTO_SORT", "1") != "0", "min_points": 1 << 18,
             "l2_table_bytes": 96 << 20, "max_table_bytes": 512 << 20}


def _auto_perm(spec, x, will_backprop=True):
    if not AUTO_SORT["enabled"] or spec.frac_mode == K._lib.HASH_REFERENCE or spec.n_levels == 0:
        return None
    nbytes = sum(spec.rows) * spec.n_feat * 4
    if x.shape[0] < AUTO_SORT["min_points"] or nbytes > AUTO_SORT["max_table_bytes"]:
        return None
    if not will_backprop and nbytes <= AUTO_SORT["l2_table_bytes"]:
        return None
    return K.morton_perm(x.detach())


class _HashEncode(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, B, spec, *tables):
        ctx.perm = _auto_perm(spec, x, will_backprop=any(ctx.needs_input_grad[3:]))
        out = K.hash_encode_fwd(spec, x, tables, B, perm=ctx.perm)
        ctx.spec = spec
        ctx.has_B = B is not None
        ctx.table_refs = tables
        ctx.save_for_backward(x, *( [B] if B is not None else [] ), *tables)
        return out[:, :spec.width]

    @staticmethod
    def backward(ctx, dy):
        spec = ctx.spec
        saved = ctx.saved_tensors
        x = saved[0]
        B = saved[1] if ctx.has_B else None
        tables = saved[2:] if ctx.has_B else saved[1:]
        need_dx = ctx.needs_input_grad[0]
        need_tab = any(ctx.needs_input_grad[3:])
        C = spec.n_fourier
        second_order = need_dx and torch.is_grad_enabled() and (dy.requires_grad or x.requires_grad)
        grads: List[Optional[torch.Tensor]] = [None] * len(tables)
        gt = None
        if need_tab:
            direct = [K.direct_grad_target(t) for t in ctx.table_refs] if getattr(ctx, "table_refs", None) else [None]
            if all(d is not None for d in direct) and all(ctx.needs_input_grad[3:]):
                gt = direct                               # the scatter accumulates straight into the tables' .grad
            else:
                flat = K.ZERO_POOL.take(sum(r * spec.n_feat for r in spec.rows), dy.device)
                gt, off = [], 0
                for r in spec.rows:
                    gt.append(flat[off:off + r * spec.n_feat].view(r, spec.n_feat))
                    off += r * spec.n_feat
                grads = list(gt)
        dx = None
        if second_order:
            # differentiable form of d/dx: only the Fourier prefix depends on x in reference mode
            if C > 0:
                if dy.dtype == torch.float32 and dy.dim() == 2 and dy.shape[1] == spec.width:
                    dx = _FourierDx.apply(x, dy, B, spec.width)
                else:
                    xp = torch.matmul(TWO_PI * x[:, :3], B)
                    dxp = dy[:, 3:3 + C] * torch.cos(xp) - dy[:, 3 + C:3 + 2 * C] * torch.sin(xp)
                    dx = dy[:, :3] + TWO_PI * torch.matmul(dxp, B.t())
            if spec.frac_mode != K._lib.HASH_REFERENCE:
                with torch.no_grad():
                    zero_pre = dy.detach().clone()
                    if C > 0:
                        zero_pre[:, :3 + 2 * C] = 0
                    hx = K.hash_encode_bwd(spec, x.detach(), tables, B, zero_pre, None, True, perm=ctx.perm)
                dx = hx if dx is None else dx + hx
            if need_tab:
                with torch.no_grad():
                    K.hash_encode_bwd(spec, x.detach(), tables, B, dy.detach(), gt, False, perm=ctx.perm) if spec.n_levels else None
        elif need_dx or need_tab:
            kernel_dx = need_dx and (C > 0 or spec.frac_mode != K._lib.HASH_REFERENCE)
            if kernel_dx or (need_tab and spec.n_levels > 0):
                with torch.no_grad():
                    dx = K.hash_encode_bwd(spec, x.detach(), tables, B, dy.detach(), gt, kernel_dx, perm=ctx.perm)
            if need_dx and dx is None:
                dx = torch.zeros_like(x[:, :3])
        if dx is not None and x.shape[1] != 3:
            pad = torch.zeros_like(x)
            pad[:, :3] = dx
            dx = pad
        return (dx, None, None, *grads)


def hash_encode(x: torch.Tensor, spec: K.HashGridSpec, tables: Sequence[torch.Tensor], B: Optional[torch.Tensor]):
    """[..., 3] -> [..., width]; last-dim layout [x | sin | cos | levels]."""
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    y = _HashEncode.apply(x2, B, spec, *tables)
    return y if len(lead) == 1 else y.reshape(*lead, spec.width)


_EMPTY_SPEC_CACHE = {}


def fourier_feature(x: torch.Tensor, B: torch.Tensor) -> torch.Tensor:
    """[x | sin(2 pi x B) | cos(2 pi x B)]  (FourierFeature.forward, frequency_enc.py:63-67)."""
    C = B.shape[1]
    spec = _EMPTY_SPEC_CACHE.get(C)
    if spec is None:
        spec = _EMPTY_SPEC_CACHE[C] = K.HashGridSpec([], [], 2, 0, C)
    return hash_encode(x, spec, (), B.contiguous())


# ---------------------------------------------------------------------------------------------
# positional encoding
# ---------------------------------------------------------------------------------------------
class _PosEncDx(torch.autograd.Function):
    """d/dx of the positional encoding on the RECORDED backward pass (ImplicitNetwork.gradient, create_graph=True, through
    the filter banks' encodings): one kernel that stays differentiable in x and dy (one more kernel) - ~9 tensor ops per
    band did it before, each with its own backward ops in the loss's backward (~1000 launches of a filter-bank step)."""

    @staticmethod
    def forward(ctx, x, dy, bands, include_input):
        ctx.bands, ctx.include_input = bands, include_input
        ctx.save_for_backward(x, dy)
        return K.posenc_bwd(x.detach(), bands, include_input, dy.detach())

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        x, dy = ctx.saved_tensors
        g_dy, g_x = K.posenc_dx_bwd(g, x, ctx.bands, ctx.include_input, dy, want_gdy=ctx.needs_input_grad[1],
                                    want_gx=ctx.needs_input_grad[0])
        return g_x, g_dy, None, None


class _PosEnc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bands, include_input):
        out = K.posenc_fwd(x, bands, include_input)
        ctx.bands, ctx.include_input = bands, include_input
        ctx.save_for_backward(x)
        return out[:, :K.posenc_width(x.shape[1], len(bands), include_input)]

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        if not ctx.needs_input_grad[0]:
            return None, None, None
        bands, inc = ctx.bands, ctx.include_input
        if torch.is_grad_enabled() and (dy.requires_grad or x.requires_grad):
            if dy.dtype == torch.float32 and x.dtype == torch.float32 and dy.dim() == 2:
                return _PosEncDx.apply(x, dy, bands, inc), None, None
            d = x.shape[1]
            base = 2 * d if inc else 0
            dx = dy[:, :d] + dy[:, d:2 * d] if inc else torch.zeros_like(x)
            for q, f in enumerate(bands):
                a = x * f
                dx = dx + f * (dy[:, base + 2 * q * d: base + (2 * q + 1) * d] * torch.cos(a)
                               - dy[:, base + (2 * q + 1) * d: base + (2 * q + 2) * d] * torch.sin(a))
            return dx, None, None
        with torch.no_grad():
            return K.posenc_bwd(x.detach(), bands, inc, dy.detach()), None, None


def positional_encoding(x: torch.Tensor, bands: Sequence[float], include_input: bool) -> torch.Tensor:
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    y = _PosEnc.apply(x2, tuple(bands), bool(include_input))
    return y if len(lead) == 1 else y.reshape(*lead, y.shape[-1])
