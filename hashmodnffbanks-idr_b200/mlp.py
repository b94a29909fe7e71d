"""MLP building blocks on top of the contraction kernel (csrc/gemm.cu).

Two execution paths share the same kernel:

* differentiable path - torch.autograd.Functions (`linear`, `linear_act`, `mm_nt/nn/tn`) whose
  backward is itself expressed with these Functions, so the reference's second-order use
  (ImplicitNetwork.gradient with create_graph=True, implicit_differentiable_renderer.py:116-128, then a
  loss on that gradient) works to any order without hand-derived double-backward kernels;
* inference path (`SdfPipeline`) - no autograd, weights folded (weight-norm) and hi/lo-split once per
  parameter version, activations handed from epilogue to next layer already split, device-side row
  count (`m_count`) so the ray tracer never syncs with the host.
"""
import math
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import kernels as K
from ._lib import IdrkError

SQRT2_INV = 1.0 / math.sqrt(2.0)

# Bumped by whoever updates parameters behind torch's back (the fused optimiser kernel writes the flat
# bucket through a raw pointer, so tensor._version does not move): invalidates folded-weight caches.
WEIGHTS_EPOCH = [0]


def weights_changed():
    WEIGHTS_EPOCH[0] += 1


# ---------------------------------------------------------------------------------------------
# operand preparation
# ---------------------------------------------------------------------------------------------
# Width of the output tiles the split-K factor of the weight-gradient (TN) contractions is sized for: the factor is chosen so
# that tiles x splits fills the SMs once (IDRK_TN_TILE: A/B knob)
TN_TILE_N = int(os.environ.get("IDRK_TN_TILE", "64"))


def _tn_split_k(M: int, N: int, Kc: int) -> int:
    if Kc < 1024 or M <= 0:
        return 1
    tiles = ((M + 127) // 128) * ((N + TN_TILE_N - 1) // TN_TILE_N)
    return max(1, min(Kc // 256, 148 // max(tiles, 1)))


def _three_pass() -> bool:
    return K.get_precision() == K.PREC_3XTF32


def _split_mode() -> Optional[str]:
    """Operand format of the autograd path's contractions: "p16" (16-bit pairs, csrc/gemm_p16.cu), "tf32" (hi / lo
    pairs, csrc/gemm.cu) or None (single pass / FFMA: no split operands)."""
    if not _three_pass():
        return None
    return "p16" if K.training_p16() else "tf32"


def _pack_key(pack):
    return "tf32" if len(pack) == 2 else int(pack[2])


def tag_split(t: torch.Tensor, *pack) -> torch.Tensor:
    """Remembers a split operand of `t` on the tensor object (valid for its current version): (hi, lo) tf32 pair or
    (h, l, fmt) 16-bit pair; one per format."""
    cur = getattr(t, "_idrk_split", None)
    ver = (t._version, WEIGHTS_EPOCH[0])       # the fused optimiser rewrites parameters without moving tensor._version
    if cur is None or cur[1] != ver:
        cur = ({}, ver)
        t._idrk_split = cur
    cur[0][_pack_key(pack)] = tuple(pack)
    return t


def split_of(t: torch.Tensor, fmt: Optional[int] = None):
    """Split operand recorded for exactly this tensor in the current operand format, or None.  16-bit pairs: `fmt`
    selects the pair format (default bf16)."""
    sp = getattr(t, "_idrk_split", None)
    mode = _split_mode()
    if sp is None or sp[1] != (t._version, WEIGHTS_EPOCH[0]) or mode is None:
        return None
    return sp[0].get("tf32" if mode == "tf32" else (K.P16_BF16 if fmt is None else fmt))


def make_split(t: torch.Tensor, fmt: Optional[int] = None):
    """Split operand of `t` for the current operand format (None when a single pass is used), cached on the tensor.
    16-bit pairs are bf16 (the full fp32 range, ~17 bits) unless the caller asks for fp16 pairs (`fmt`; ~22 bits,
    |x| < 65504 - the SIREN layers of the filter banks, whose sin(w0 .) chains amplify operand rounding)."""
    mode = _split_mode()
    if mode is None:
        return None
    sp = split_of(t, fmt)
    if sp is not None:
        return sp
    if mode == "p16":
        sp = K.split_p16(t.detach(), K.P16_BF16 if fmt is None else fmt)
    else:
        sp = K.split_tf32(K.operand(t.detach()))
    tag_split(t, *sp)
    return sp


def _prep(t: torch.Tensor, split=None):
    """(main, lo) operand pair for the current precision (tf32 / fp32 kernels)."""
    if _three_pass():
        if split is None:
            split = split_of(t)
        if split is not None:
            return split
        return K.split_tf32(K.operand(t.detach()))
    return K.operand(t.detach()), None


def _pack_for(t: torch.Tensor, pack, fmt: int):
    """16-bit pair of `t` in `fmt`: the caller's pack when it has that format, else the tensor's cached / a fresh split."""
    if pack is not None and len(pack) == 3 and pack[2] == fmt:
        return pack
    return make_split(t, fmt)


def _act_bwd_tagged(dH, dS, S, H, mode: int, act: float, scale: float, want_split: bool = True) -> torch.Tensor:
    """dZ of the fused activation backward, tagged with its split operand (cotangent: bf16 pair in the 16-bit format)."""
    sm = _split_mode() if want_split else None
    if sm == "p16":
        dZ, pack = K.act_bwd_p16(dH, dS, S, H, mode, act, scale, K.P16_BF16)
        return tag_split(dZ, *pack)
    dZ, hi, lo = K.act_bwd(dH, dS, S, H, mode, act, scale, sm == "tf32")
    if hi is not None:
        tag_split(dZ, hi, lo)
    return dZ


def _raw_mm(layout: int, A: torch.Tensor, B: torch.Tensor, M: int, N: int, Kc: int, bias=None,
            mode=K.EPI_NONE, act=0.0, scale=1.0, want_s=False, a_split=None, b_split=None, split_out=False, op_fmt=None):
    dev = A.device
    split_k = _tn_split_k(M, N, Kc) if layout == K.GEMM_TN else 1
    if split_k > 1:                                   # split-K accumulates with atomics: zero-initialised output
        C = K.ZERO_POOL.take(M * K.pad4(N), dev).view(M, K.pad4(N))[:, :N]
    else:
        C = K.empty_padded(M, N, dev)
    S = K.empty_padded(M, N, dev) if want_s else None
    if M == 0:
        return C, S
    if _split_mode() == "p16":
        # one launch cannot mix fp16 and bf16 pairs: everything is a bf16 pair unless the caller asks for fp16 (`op_fmt`)
        fmt = K.P16_BF16 if op_fmt is None else op_fmt
        a, b = _pack_for(A, a_split, fmt), _pack_for(B, b_split, fmt)
        C_pair = None
        if split_out and split_k == 1:
            C_pair = (*K.empty_pair16(M, N, dev, fmt), fmt)
        K.gemm_p16(layout, a, b, M, N, Kc, C=C, C_pair=C_pair, S=S, bias=bias, mode=mode, act=act, scale=scale,
                   split_k=split_k)
        if C_pair is not None:
            tag_split(C, *C_pair)
        return C, S
    if a_split is not None and len(a_split) == 3:
        a_split = None
    if b_split is not None and len(b_split) == 3:
        b_split = None
    a, a_lo = _prep(A, a_split)
    b, b_lo = _prep(B, b_split)
    C_hi = C_lo = None
    if split_out and _three_pass() and split_k == 1:
        C_hi, C_lo = K.empty_padded(M, N, dev), K.empty_padded(M, N, dev)
    K.gemm(layout, a, b, M, N, Kc, A_lo=a_lo, B_lo=b_lo, C=C, C_hi=C_hi, C_lo=C_lo, S=S, bias=bias, mode=mode, act=act,
           scale=scale, split_k=split_k)
    if C_hi is not None:
        tag_split(C, C_hi, C_lo)
    return C, S


# ---------------------------------------------------------------------------------------------
# differentiable contractions (closed under differentiation)
# ---------------------------------------------------------------------------------------------
class _MMNT(torch.autograd.Function):
    """C[M,N] = A[M,K] @ B[N,K]^T"""

    @staticmethod
    def forward(ctx, A, B, a_split, b_split):
        ctx.a_split, ctx.b_split = (a_split or split_of(A)), (b_split or split_of(B))
        ctx.save_for_backward(A, B)
        return _raw_mm(K.GEMM_NT, A, B, A.shape[0], B.shape[0], A.shape[1], a_split=ctx.a_split, b_split=ctx.b_split)[0]

    @staticmethod
    def backward(ctx, dC):
        A, B = ctx.saved_tensors
        dsp = make_split(dC) if (ctx.needs_input_grad[0] and ctx.needs_input_grad[1]) else None
        dA = mm_nn(dC, B, dsp, ctx.b_split) if ctx.needs_input_grad[0] else None
        dB = mm_tn(dC, A, dsp, ctx.a_split) if ctx.needs_input_grad[1] else None
        return dA, dB, None, None


class _MMNN(torch.autograd.Function):
    """C[M,N] = A[M,K] @ B[K,N]"""

    @staticmethod
    def forward(ctx, A, B, a_split, b_split):
        ctx.a_split, ctx.b_split = (a_split or split_of(A)), (b_split or split_of(B))
        ctx.b_ref = getattr(B, "_idrk_wn", None)
        ctx.save_for_backward(A, B)
        return _raw_mm(K.GEMM_NN, A, B, A.shape[0], B.shape[1], A.shape[1], a_split=ctx.a_split, b_split=ctx.b_split)[0]

    @staticmethod
    def backward(ctx, dC):
        A, B = ctx.saved_tensors
        ent = None
        if ctx.needs_input_grad[1] and dC.shape[0] > 0:
            # B is a layer weight inside the recorded backward (dX = dZ W): its gradient A^T dC joins the weight's
            # accumulation buffer on the side stream (mlp._LeafSide) instead of going through autograd's adds
            ent = _leaf_entry(ctx.b_ref, B.shape, dC.device)
        dsp = make_split(dC) if ((ctx.needs_input_grad[0] and ctx.needs_input_grad[1]) or ent is not None) else None
        if ent is not None:
            asp = _pack_for(A, ctx.a_split, K.P16_BF16) if _split_mode() == "p16" else (ctx.a_split or make_split(A))
            Ad, dCd = A.detach(), dC.detach()
            with LEAF_SIDE.fork():
                _tn_accumulate(Ad, asp, dCd, dsp, ent[2])
            LEAF_SIDE.keep.append((A, asp, dC, dsp))
            dA = mm_nt(dC, B, dsp, ctx.b_split) if ctx.needs_input_grad[0] else None
            return dA, None, None, None
        dA = mm_nt(dC, B, dsp, ctx.b_split) if ctx.needs_input_grad[0] else None
        dB = mm_tn(A, dC, ctx.a_split, dsp) if ctx.needs_input_grad[1] else None
        return dA, dB, None, None


class _MMTN(torch.autograd.Function):
    """C[M,N] = A[K,M]^T @ B[K,N]   (contraction over rows: weight gradients)"""

    @staticmethod
    def forward(ctx, A, B, a_split, b_split):
        ctx.a_split, ctx.b_split = (a_split or split_of(A)), (b_split or split_of(B))
        ctx.save_for_backward(A, B)
        return _raw_mm(K.GEMM_TN, A, B, A.shape[1], B.shape[1], A.shape[0], a_split=ctx.a_split, b_split=ctx.b_split)[0]

    @staticmethod
    def backward(ctx, dC):
        A, B = ctx.saved_tensors
        dsp = make_split(dC) if (ctx.needs_input_grad[0] and ctx.needs_input_grad[1]) else None
        dA = mm_nt(B, dC, ctx.b_split, dsp) if ctx.needs_input_grad[0] else None
        dB = mm_nn(A, dC, ctx.a_split, dsp) if ctx.needs_input_grad[1] else None
        return dA, dB, None, None


def mm_nt(A, B, a_split=None, b_split=None):
    return _MMNT.apply(A, B, a_split, b_split)


def mm_nn(A, B, a_split=None, b_split=None):
    return _MMNN.apply(A, B, a_split, b_split)


def mm_tn(A, B, a_split=None, b_split=None):
    return _MMTN.apply(A, B, a_split, b_split)


class _ColSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.rows = x.shape[0]
        return K.colsum(x.detach())

    @staticmethod
    def backward(ctx, g):
        return g.unsqueeze(0).expand(ctx.rows, -1)


def colsum(x):
    return _ColSum.apply(x)


_direct_target = K.direct_grad_target


class _LeafSide:
    """Leaf-gradient work of a trainer step on a SIDE stream.

    The backward pass of a step is one long chain of small dependent launches (dZ -> dX -> dZ -> ...); the weight-gradient
    contractions (dW = dZ^T X), the bias column sums and the weight-norm backward only feed the optimiser, so they do not
    belong on that chain.  Between `begin()` and `finish()` (DataParallelTrainer brackets loss.backward() with them)
    every weight-normalised layer's dW is ACCUMULATED by the TN contraction (atomic split-K epilogue) into one zeroed
    buffer per weight - all uses of a shared weight land in the same buffer, so autograd's add kernels disappear - on the
    side stream, autograd gets None for it, and `finish()` runs the weight-norm backward from those buffers straight into
    the flat gradient bucket and joins the streams.  Inside a CUDA-graph capture the fork / join become graph edges, i.e.
    a parallel branch.  Operands the side stream reads are kept referenced until the join (no allocator reuse races).
    Only used when gradients are not being recorded (first-order backward)."""

    def __init__(self):
        self.active = False
        self.stream = None
        self.keep: List = []
        self.acc: Dict = {}
        self.enabled = os.environ.get("IDRK_LEAF_SIDE", "1") != "0"       # A/B switch

    def begin(self, device):
        if not self.enabled:
            return
        if self.stream is None or self.stream.device != torch.device(device):
            self.stream = torch.cuda.Stream(device)
        self.active, self.keep, self.acc = True, [], {}
        self.forked = False

    def fork(self):
        self.stream.wait_stream(torch.cuda.current_stream())
        self.forked = True
        return torch.cuda.stream(self.stream)

    def finish(self):
        if not self.active:
            return
        try:
            if self.acc:
                with torch.no_grad(), self.fork():
                    for g, v, buf, tg, tv in self.acc.values():
                        K.weight_norm_bwd(g, v, buf, into=(tg, tv))
            if self.forked:
                torch.cuda.current_stream().wait_stream(self.stream)
        finally:
            self.active, self.keep, self.acc = False, [], {}


LEAF_SIDE = _LeafSide()


def _leaf_entry(wn, shape, device):
    """Accumulation buffer of a weight-normalised weight W = g v / ||v|| (wn = (g, v)) for this step, or None when the
    side path does not apply (not a trainer step, gradients being recorded, no direct gradient targets)."""
    if not LEAF_SIDE.active or wn is None or torch.is_grad_enabled() or _split_mode() is None:
        return None
    ent = LEAF_SIDE.acc.get(id(wn[1]))
    if ent is None:
        tg, tv = _direct_target(wn[0]), _direct_target(wn[1])
        if tg is None or tv is None:
            return None
        n_out, n_in = shape
        buf = K.ZERO_POOL.take(n_out * K.pad4(n_in), device).view(n_out, K.pad4(n_in))[:, :n_in]
        ent = LEAF_SIDE.acc[id(wn[1])] = (wn[0], wn[1], buf, tg, tv)
    return ent


def _tn_accumulate(dZ, dsp, X, xsp, buf):
    """buf[out, in] += dZ^T X on the current stream (split-K with atomic accumulation; `buf` zero-initialised)."""
    M, N, Kc = dZ.shape[1], X.shape[1], dZ.shape[0]
    split_k = _tn_split_k(M, N, Kc)
    if _split_mode() == "p16":
        K.gemm_p16(K.GEMM_TN, dsp, xsp, M, N, Kc, C=buf, accumulate=True, split_k=split_k)
    else:
        K.gemm(K.GEMM_TN, dsp[0], xsp[0], M, N, Kc, A_lo=dsp[1], B_lo=xsp[1], C=buf, accumulate=True, split_k=split_k)


def _layer_backward(ctx, dZ, X, W):
    """Shared backward of Z = X W^T + b: one hi/lo split of dZ feeds both contractions."""
    need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
    need_b = ctx.has_bias and ctx.needs_input_grad[2]
    if getattr(ctx, "w_ref", None) is not None:
        W._idrk_wn = ctx.w_ref            # the unpacked saved tensor is a new object: keep the weight's identity for mm_nn
    b_tgt = _direct_target(getattr(ctx, "bias_ref", None)) if need_b else None
    if b_tgt is not None and not (b_tgt.dim() == 1 and b_tgt.numel() == dZ.shape[1]):
        b_tgt = None
    ent = None
    if need_w and dZ.shape[0] > 0 and (not need_b or b_tgt is not None):
        ent = _leaf_entry(getattr(ctx, "w_ref", None), W.shape, dZ.device)
    dsp = make_split(dZ) if ((need_x and need_w) or ent is not None) else None
    if ent is not None:
        xsp = _pack_for(X, ctx.x_split, K.P16_BF16) if _split_mode() == "p16" else (ctx.x_split or make_split(X))
        dZd = dZ.detach()
        with LEAF_SIDE.fork():
            _tn_accumulate(dZd, dsp, X, xsp, ent[2])
            if need_b:
                K.colsum(dZd, into=b_tgt)
        LEAF_SIDE.keep.append((dZ, dsp, X, xsp))
        dX = mm_nn(dZ, W, dsp, ctx.w_split) if need_x else None
        return dX, None, None
    dX = mm_nn(dZ, W, dsp, ctx.w_split) if need_x else None
    dW = mm_tn(dZ, X, dsp, ctx.x_split) if need_w else None
    db = None
    if need_b:
        if b_tgt is not None:
            K.colsum(dZ.detach(), into=b_tgt)
        else:
            db = colsum(dZ)
    return dX, dW, db


class _Linear(torch.autograd.Function):
    """Z = X W^T + b  (bias fused in the epilogue)."""

    @staticmethod
    def forward(ctx, X, W, b):
        ctx.x_split, ctx.w_split = make_split(X), make_split(W)
        ctx.save_for_backward(X, W)
        ctx.has_bias = b is not None
        ctx.bias_ref = b
        ctx.w_ref = getattr(W, "_idrk_wn", None)
        return _raw_mm(K.GEMM_NT, X, W, X.shape[0], W.shape[0], X.shape[1], bias=b.detach() if b is not None else None,
                       a_split=ctx.x_split, b_split=ctx.w_split)[0]

    @staticmethod
    def backward(ctx, dZ):
        X, W = ctx.saved_tensors
        return _layer_backward(ctx, dZ, X, W)


def linear(X, W, b=None):
    return _Linear.apply(X, W, b)


_ACT_MODES = {"softplus": K.EPI_SOFTPLUS, "relu": K.EPI_RELU, "sine": K.EPI_SINE, "tanh": K.EPI_TANH}

# Parity-test tap: when set to a list, every ReLU layer appends its derivative mask S = 1[z > 0] (the activation pattern).
# ReLU gradients are discontinuous in z, so a checker has to compare them under the SAME pattern (tests/test_full_shapes.py).
ACT_PATTERN_TAP = [None]


class _MulAct(torch.autograd.Function):
    """dZ = scale * dH * S on the RECORDED backward pass (ImplicitNetwork.gradient, create_graph=True): one kernel forms the
    product and its 3xTF32 operand pair (a torch mul + a split launch before), and stays differentiable in dH and S -
    the loss on the gradient (eikonal term, normals) back-propagates through it with the same kernel."""

    @staticmethod
    def forward(ctx, dH, S, scale):
        ctx.scale = scale
        ctx.save_for_backward(dH, S)
        return _act_bwd_tagged(dH.detach(), None, S.detach(), None, K.EPI_RELU, 0.0, scale)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        dH, S = ctx.saved_tensors
        g_dH = g_S = None
        if ctx.needs_input_grad[0]:
            g_dH = _act_bwd_tagged(g, None, S, None, K.EPI_RELU, 0.0, ctx.scale)
        if ctx.needs_input_grad[1]:
            g_S = _act_bwd_tagged(g, None, dH, None, K.EPI_RELU, 0.0, ctx.scale, want_split=False)
        return g_dH, g_S, None


class _LinearAct(torch.autograd.Function):
    """(H, S) = act(X W^T + b) and its derivative, both from one kernel epilogue.

    H = scale * act(Z),  S = act'(Z).  S is returned as a differentiable output so that when the
    backward pass is itself recorded (create_graph) the dependence of act' on Z is tracked:
    d S / d Z = act''(Z), written below per activation in terms of the saved outputs.
    In 3xTF32 mode the epilogue also writes H pre-split (hi/lo) for the next layer's contraction."""

    @staticmethod
    def forward(ctx, X, W, b, mode, act, scale):
        # SIREN layers (filter banks): sin(w0 z) multiplies operand rounding by w0, so their forward products use
        # fp16 pairs (inputs and sine outputs are bounded); their backward products re-split to bf16 on demand
        fmt = K.P16_FP16 if (mode == "sine" and _split_mode() == "p16") else None
        ctx.x_split, ctx.w_split = make_split(X, fmt), make_split(W, fmt)
        H, S = _raw_mm(K.GEMM_NT, X, W, X.shape[0], W.shape[0], X.shape[1],
                       bias=b.detach() if b is not None else None, mode=_ACT_MODES[mode], act=act, scale=scale,
                       want_s=True, a_split=ctx.x_split, b_split=ctx.w_split, split_out=True, op_fmt=fmt)
        ctx.mode, ctx.act, ctx.scale, ctx.has_bias = mode, act, scale, b is not None
        ctx.bias_ref = b
        ctx.w_ref = getattr(W, "_idrk_wn", None)
        if ACT_PATTERN_TAP[0] is not None and mode == "relu":
            ACT_PATTERN_TAP[0].append(S.detach().clone())
        ctx.save_for_backward(X, W, H, S)
        ctx.set_materialize_grads(False)
        return H, S

    @staticmethod
    def backward(ctx, dH, dS):
        X, W, H, S = ctx.saved_tensors
        if dH is None and dS is None:
            return None, None, None, None, None, None
        if not torch.is_grad_enabled():
            # plain (first-order) backward: one fused kernel forms dZ and its 3xTF32 operand pair
            dZ = _act_bwd_tagged(dH, dS if ctx.mode != "relu" else None, S, H, _ACT_MODES[ctx.mode], ctx.act,
                                 ctx.scale) if (dH is not None or ctx.mode != "relu") else None
            if dZ is None:
                return None, None, None, None, None, None
            return (*_layer_backward(ctx, dZ, X, W), None, None, None)
        dZ = None
        if dH is not None:
            if dS is None and dH.dim() == 2 and dH.dtype == torch.float32:
                dZ = _MulAct.apply(dH, S, ctx.scale)
            else:
                dZ = dH * S if ctx.scale == 1.0 else dH * (S * ctx.scale)
        if dS is not None:
            if ctx.mode == "softplus":
                s2 = (ctx.act * S) * (1.0 - S)
            elif ctx.mode == "sine":
                s2 = H * (-(ctx.act ** 2) / ctx.scale)
            elif ctx.mode == "tanh":
                s2 = (H * S) * (-2.0 / ctx.scale)
            else:
                s2 = None
            if s2 is not None:
                dZ = dS * s2 if dZ is None else dZ + dS * s2
        if dZ is None:
            return None, None, None, None, None, None
        return (*_layer_backward(ctx, dZ, X, W), None, None, None)


def linear_act(X, W, b, mode: str, act: float = 0.0, scale: float = 1.0):
    """scale * act(X W^T + b); activations: softplus(beta=act), relu, sine(w0=act), tanh."""
    return _LinearAct.apply(X, W, b, mode, float(act), float(scale))[0]


class _SquashRows(torch.autograd.Function):
    """out = x with column 0 -> tanh(s / (2 + rho(s))), rho a constant for autograd (the reference's last step of
    ImplicitNetwork.forward, implicit_differentiable_renderer.py:108-113 - a dozen tensor ops there, and as many again in
    each recorded backward).  One kernel gives out, d = d out_0 / d s and d2 = d^2 out_0 / d s^2; d is returned as a
    differentiable OUTPUT so that the recorded backward (`gradient()`, create_graph) keeps its dependence on s:
    d d / d s = d2.  Third order is never needed (d2 is a constant)."""

    @staticmethod
    def forward(ctx, x, beta):
        out, d, d2 = K.sdf_squash_rows(x.detach(), beta, True)
        ctx.cols = x.shape[1]
        ctx.save_for_backward(d, d2)
        ctx.set_materialize_grads(False)
        return out, d

    @staticmethod
    def backward(ctx, g_out, g_d):
        d, d2 = ctx.saved_tensors
        gx = None
        if g_out is not None:
            gx = torch.cat([g_out[:, :1] * d, g_out[:, 1:]], 1)
        if g_d is not None:
            extra = torch.nn.functional.pad(g_d * d2, (0, ctx.cols - 1))
            gx = extra if gx is None else gx + extra
        return gx, None


def sdf_squash_rows(x: torch.Tensor, beta: float) -> torch.Tensor:
    return _SquashRows.apply(x, float(beta))[0]


class _WeightNorm(torch.autograd.Function):
    """W = g * v / ||v||_row  (legacy nn.utils.weight_norm, dim=0)."""

    @staticmethod
    def forward(ctx, g, v):
        ctx.save_for_backward(g, v)
        ctx.set_materialize_grads(False)        # every use may have routed its dW through mlp._LeafSide: then dW is None
        sm = _split_mode()
        if sm == "p16":
            W, pack = K.weight_norm_fwd_p16(g.detach(), v.detach(), K.P16_BF16)
            tag_split(W, *pack)
        else:
            out = K.weight_norm_fwd(g.detach(), v.detach(), sm == "tf32", False)
            W = out["W"]
            if sm == "tf32":
                tag_split(W, out["W_hi"], out["W_lo"])
        W._idrk_wn = (g, v)                 # lets a trainer step route this weight's gradient work off the main chain (_LeafSide)
        return W

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dW):
        if dW is None:
            return None, None
        g, v = ctx.saved_tensors
        tg, tv = _direct_target(g), _direct_target(v)
        if tg is not None and tv is not None and ctx.needs_input_grad[0] and ctx.needs_input_grad[1]:
            K.weight_norm_bwd(g, v, dW, into=(tg, tv))
            return None, None
        dg, dv = K.weight_norm_bwd(g, v, dW)
        return dg, dv


def weight_norm(g, v):
    return _WeightNorm.apply(g, v)


_WEIGHT_SCOPE = []      # stack of per-step caches opened by `shared_weights()`


class shared_weights:
    """Within this context every layer's effective weight W = g v / ||v|| is computed ONCE and shared by all
    forward calls (IDRNetwork.shade evaluates the implicit network three times per step): autograd sums the
    weight gradients of all uses and runs the weight-norm backward once."""

    def __enter__(self):
        _WEIGHT_SCOPE.append({})
        return self

    def __exit__(self, *exc):
        _WEIGHT_SCOPE.pop()
        return False


def layer_weight(lin: torch.nn.Module) -> torch.Tensor:
    """Effective weight of a (possibly weight-normalised) nn.Linear, differentiable."""
    if not hasattr(lin, "weight_g"):
        return lin.weight
    if _WEIGHT_SCOPE and torch.is_grad_enabled():
        cache = _WEIGHT_SCOPE[-1]
        W = cache.get(id(lin))
        if W is None:
            W = cache[id(lin)] = weight_norm(lin.weight_g, lin.weight_v)
        return W
    return weight_norm(lin.weight_g, lin.weight_v)


# ---------------------------------------------------------------------------------------------
# inference pipeline (no autograd)
# ---------------------------------------------------------------------------------------------
class FoldedLayer:
    """Effective (weight-norm folded) weight of one Linear, optionally hi/lo split, in persistent buffers."""
    __slots__ = ("lin", "split", "bufs", "W", "W_lo", "Wfull", "bias", "n_out", "n_in", "W_h16", "W_l16")

    def __init__(self, lin: torch.nn.Module, split: bool):
        self.lin, self.split, self.bufs = lin, split, None
        self.W_h16 = self.W_l16 = None
        self.n_out, self.n_in = (lin.weight_v if hasattr(lin, "weight_v") else lin.weight).shape
        self.refresh()

    @torch.no_grad()
    def refresh(self):
        lin = self.lin
        if hasattr(lin, "weight_g"):
            self.bufs = K.weight_norm_fwd(lin.weight_g, lin.weight_v, self.split, False, out=self.bufs)
        else:
            self.bufs = K.weight_norm_fwd(None, lin.weight, self.split, False, out=self.bufs)
        self.W = self.bufs["W_hi"] if self.split else self.bufs["W"]
        self.W_lo = self.bufs.get("W_lo")
        self.Wfull = self.bufs["W"]
        self.bias = lin.bias.detach()
        if K.inference_fp16x2():                    # fp16 pair of the folded weight (persistent buffers)
            if self.W_h16 is None:
                self.W_h16 = K.empty_half(self.n_out, self.n_in, self.Wfull.device)
                self.W_l16 = K.empty_half(self.n_out, self.n_in, self.Wfull.device)
            K.split_f16_into(self.Wfull, self.n_out, self.n_in, 1.0, self.W_h16, self.W_l16, K.pad8(self.n_in),
                             K.pad8(self.n_in) - self.n_in)


def params_version(params: Sequence[torch.Tensor]) -> Tuple:
    return (WEIGHTS_EPOCH[0],) + tuple((p.data_ptr(), p._version) for p in params)


class SdfPipeline:
    """ImplicitNetwork forward without autograd, as the ray tracer and the eval consumers use it
    (implicit_differentiable_renderer.py:89-113 under no_grad; ray_tracing.py calls it ~50x per step).

    * weights: weight-norm folded and hi/lo-split once per parameter version;
    * activations: written by each layer's epilogue directly as the next layer's (split) operand;
      the skip concat is realised by letting layer (skip-1) write scale 1/sqrt(2) into the left
      columns of a shared buffer and copying the scaled embedding into the right columns;
    * `m_count` (device int32): number of valid rows, so compacted ray sets need no host sync;
    * want="sdf": the last Linear is reduced to its row 0 + the Laplace squash (the only column
      the tracer consumes, 6.7 % fewer MACs);  want="full": all 1 + feature columns.
    """

    def __init__(self, net):
        self.net = net
        self.n_lin = net.num_layers - 1
        self._folded = None
        self._version = None
        self._precision = None
        self._bufs: Dict = {}
        self._beta = None
        self._beta_version = None

    # -- cached state --------------------------------------------------------------------
    def layers(self):
        return [getattr(self.net, "lin%d" % l) for l in range(self.n_lin)]

    def folded(self, force: bool = False) -> List["FoldedLayer"]:
        params = [p for l in self.layers() for p in l.parameters()]
        ver = params_version(params)
        prec = K.get_precision()
        if self._folded is None or prec != self._precision:
            self._folded = [FoldedLayer(l, prec == K.PREC_3XTF32) for l in self.layers()]
        elif force or ver != self._version:
            for f in self._folded:
                f.refresh()                    # in place: buffer addresses captured in CUDA graphs stay valid
        self._version, self._precision = ver, prec
        return self._folded

    def beta(self) -> float:
        p = self.net.dencity_net.beta
        ver = (p.data_ptr(), p._version, WEIGHTS_EPOCH[0] if p.requires_grad and p.grad is not None else 0)
        if self._beta is None or ver != self._beta_version:
            self._beta = abs(float(p.detach().cpu())) + 1e-4
            self._beta_version = ver
        return self._beta

    def _buf(self, name, rows, cols, device):
        key = (name, cols)
        b = self._bufs.get(key)
        if b is None or b.shape[0] < rows or b.device != device:
            K.retire_scratch(b)              # CUDA graphs may hold its address: parked, never freed (kernels.py)
            b = torch.empty((max(rows, 1), K.pad4(cols)), device=device, dtype=torch.float32)
            self._bufs[key] = b
        return b

    def _hbuf(self, name, rows, cols, device):
        key = ("h16", name, cols)
        b = self._bufs.get(key)
        if b is None or b.shape[0] < rows or b.device != device:
            K.retire_scratch(b)
            b = torch.empty((max(rows, 1), K.pad8(cols)), device=device, dtype=torch.float16)
            self._bufs[key] = b
        return b

    def _run_f16(self, emb, rows, want, m_count, out, fl, encode=None):
        """Same pipeline with fp16-pair operands (csrc/gemm.cu gemm_f16s_kernel): half the operand bytes and half the
        tensor-pipe time of the 3xTF32 pair at the same accuracy class; used only here (no autograd)."""
        net = self.net
        dev = emb.device if emb is not None else fl[0].Wfull.device
        E = fl[0].n_in
        cur_h, cur_l = self._hbuf("emb_h", rows, E, dev), self._hbuf("emb_l", rows, E, dev)
        # the layer feeding the skip connection writes into a dedicated buffer whose last E columns (the embedding's
        # 1/sqrt(2) copy) are filled here, by the same launch that splits the embedding
        second = None
        skip_bufs = None
        for l, f in enumerate(fl[:-1]):
            if (l + 1) in net.skip_in:
                width = f.n_out + E
                skip_bufs = (l, self._hbuf(("hs", l), rows, width, dev), self._hbuf(("ls", l), rows, width, dev))
                ldw = K.pad8(width)
                second = (skip_bufs[1][:, f.n_out:], skip_bufs[2][:, f.n_out:], ldw, ldw - width, SQRT2_INV)
                break
        if encode is not None:
            # the encoder writes the operand pair itself (one launch instead of encode + split)
            encode(cur_h, cur_l, K.pad8(E), K.pad8(E) - E, second)
        else:
            K.split_f16_into(emb, rows, E, 1.0, cur_h, cur_l, K.pad8(E), K.pad8(E) - E, m_count, second=second)
        cur_dim = E
        n = self.n_lin
        for l, f in enumerate(fl):
            last = l == n - 1
            if last:
                C = self._buf("full_out", rows, f.n_out, dev) if out is None else out
                K.gemm_f16s(cur_h, cur_l, f.W_h16, f.W_l16, rows, f.n_out, cur_dim, C=C, bias=f.bias, m_count=m_count)
                res = C[:rows, :f.n_out]
                sq, _ = K.sdf_squash(res[:, 0].contiguous(), self.beta(), False)
                res[:, 0] = sq
                return res
            feeds_skip = (l + 1) in net.skip_in
            width = f.n_out + (E if feeds_skip else 0)
            head_next = (l == n - 2) and want == "sdf"
            scale = SQRT2_INV if feeds_skip else 1.0
            if head_next:
                # last hidden layer + SDF head in one launch: the epilogue multiplies the fp32 activations with row 0
                # of the last Linear and leaves one partial per 32-column group; the 512-column activation (67 MB
                # per 32 K-point chunk, written and read back before) never exists.  The head kernel then only
                # sums the partials in a fixed order, adds the bias and applies the Laplace squash.
                groups = (f.n_out + 31) // 32
                part = self._buf("sdf_partials", rows, groups, dev)
                ones = self._bufs.get("ones")
                if ones is None or ones.numel() < groups or ones.device != dev:
                    ones = self._bufs["ones"] = torch.ones(max(groups, 32), device=dev, dtype=torch.float32)
                K.gemm_f16s(cur_h, cur_l, f.W_h16, f.W_l16, rows, f.n_out, cur_dim, bias=f.bias, mode=K.EPI_SOFTPLUS,
                            act=100.0, scale=scale, m_count=m_count, dot_w=fl[n - 1].Wfull[0], dot_out=part)
                res = out if out is not None else torch.empty(rows, device=dev, dtype=torch.float32)
                K.sdf_head(part, ones[:groups], fl[n - 1].bias, self.beta(), res, rows, m_count)
                return res
            if feeds_skip and skip_bufs is not None and skip_bufs[0] == l:
                nxt_h, nxt_l = skip_bufs[1], skip_bufs[2]
            else:
                nxt_h, nxt_l = self._hbuf(("h", l & 1), rows, width, dev), self._hbuf(("l", l & 1), rows, width, dev)
            K.gemm_f16s(cur_h, cur_l, f.W_h16, f.W_l16, rows, f.n_out, cur_dim, C_h=nxt_h, C_l=nxt_l, bias=f.bias,
                        mode=K.EPI_SOFTPLUS, act=100.0, scale=scale, m_count=m_count)
            if feeds_skip and not (skip_bufs is not None and skip_bufs[0] == l):
                ldw = K.pad8(width)
                K.split_f16_into(emb, rows, E, SQRT2_INV, nxt_h[:, f.n_out:], nxt_l[:, f.n_out:], ldw, ldw - width, m_count)
            cur_h, cur_l, cur_dim = nxt_h, nxt_l, width
        raise IdrkError("unreachable")

    # -- execution -----------------------------------------------------------------------
    @torch.no_grad()
    def run(self, emb: Optional[torch.Tensor], rows: int, want: str = "sdf", m_count: Optional[torch.Tensor] = None,
            out: Optional[torch.Tensor] = None, encode=None) -> torch.Tensor:
        """emb: fp32 embedding buffer [>=rows, ld] (padded operand), width = net input width.  `encode(h, l, ld, pad,
        second)` (fp16-pair mode only): a callable that writes the embedding's operand pair itself; emb may be None."""
        net = self.net
        fl = self.folded()
        if K.inference_fp16x2():
            if fl[0].W_h16 is None:
                fl = self.folded(force=True)
            return self._run_f16(emb, rows, want, m_count, out, fl, encode)
        if emb is None:
            raise IdrkError("SdfPipeline.run: an fp32 embedding is required outside the fp16-pair mode")
        split = K.get_precision() == K.PREC_3XTF32
        dev = emb.device
        E = fl[0].n_in
        if split:
            cur = self._buf("emb_hi", rows, E, dev)
            cur_lo = self._buf("emb_lo", rows, E, dev)
            K.split_into(emb, rows, E, 1.0, cur, cur_lo, K.pad4(E), K.pad4(E) - E, m_count)
        else:
            cur, cur_lo = emb, None
        cur_dim = E
        n = self.n_lin
        for l, f in enumerate(fl):
            last = l == n - 1
            if last:
                if want == "sdf":
                    res = out if out is not None else torch.empty(rows, device=dev, dtype=torch.float32)
                    K.sdf_head(cur, f.Wfull[0], f.bias, self.beta(), res, rows, m_count)
                    return res
                C = self._buf("full_out", rows, f.n_out, dev) if out is None else out
                K.gemm(K.GEMM_NT, cur, f.W, rows, f.n_out, cur_dim, A_lo=cur_lo, B_lo=f.W_lo, C=C, bias=f.bias,
                       m_count=m_count)
                res = C[:rows, :f.n_out]
                sq, _ = K.sdf_squash(res[:, 0].contiguous(), self.beta(), False)
                res[:, 0] = sq
                return res
            feeds_skip = (l + 1) in net.skip_in
            width = f.n_out + (E if feeds_skip else 0)
            head_next = (l == n - 2) and want == "sdf"           # next consumer is the fp32 SDF head
            use_split = split and not head_next
            nxt = self._buf(("h", l & 1), rows, width, dev)
            nxt_lo = self._buf(("l", l & 1), rows, width, dev) if use_split else None
            K.gemm(K.GEMM_NT, cur, f.W, rows, f.n_out, cur_dim, A_lo=cur_lo, B_lo=f.W_lo,
                   C=None if use_split else nxt, C_hi=nxt if use_split else None, C_lo=nxt_lo, bias=f.bias,
                   mode=K.EPI_SOFTPLUS, act=100.0, scale=SQRT2_INV if feeds_skip else 1.0, m_count=m_count)
            if feeds_skip:
                ldw = K.pad4(width)
                K.split_into(emb, rows, E, SQRT2_INV, nxt[:, f.n_out:], nxt_lo[:, f.n_out:] if use_split else None,
                             ldw, ldw - width, m_count)
            cur, cur_lo, cur_dim = nxt, nxt_lo, width
        raise IdrkError("unreachable")
