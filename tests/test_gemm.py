"""GPU checks of the MLP contraction tiles (tcgen05/TMEM/TMA kernel and the FFMA kernel) against
fp64 matmul.  Tolerances relative to sum|a||b| per output: fp32 1e-6, 3xTF32 1e-5, TF32 2e-3."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"

TOL = {"fp32": 2e-6, "3xtf32": 1e-5, "tf32": 3e-3}


def _mk(shape, seed):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)).to(DEV)


def _run(layout, prec, M, N, Kc, **kw):
    from idrk import kernels as K
    A = _mk((M, Kc), 1) if layout != K.GEMM_TN else _mk((Kc, M), 1)
    B = _mk((N, Kc), 2) if layout == K.GEMM_NT else _mk((Kc, N), 2)
    Ao, Bo = K.operand(A), K.operand(B)
    p = K._PRECISION[prec]
    A_lo = B_lo = None
    Au, Bu = Ao, Bo
    if p == K.PREC_3XTF32:
        Au, A_lo = K.split_tf32(Ao)
        Bu, B_lo = K.split_tf32(Bo)
    C = K.empty_padded(M, N, DEV)
    C.fill_(float("nan"))
    K.gemm(layout, Au, Bu, M, N, Kc, precision=p, A_lo=A_lo, B_lo=B_lo, C=C, **kw)
    A64 = A.double() if layout != K.GEMM_TN else A.double().t()
    B64 = B.double() if layout == K.GEMM_NT else B.double().t()
    ref = A64 @ B64.t()
    mag = A64.abs() @ B64.abs().t()
    return C, ref, mag


@pytest.mark.parametrize("prec", ["fp32", "tf32", "3xtf32"])
@pytest.mark.parametrize("layout", [0, 1, 2])
@pytest.mark.parametrize("shape", [(128, 128, 32), (300, 445, 67), (2048, 512, 512), (77, 3, 512), (1000, 257, 40)])
def test_gemm_plain(layout, prec, shape):
    M, N, Kc = shape
    C, ref, mag = _run(layout, prec, M, N, Kc)
    err = ((C.double() - ref).abs() / (mag + 1e-30)).max().item()
    assert err < TOL[prec], err


@pytest.mark.parametrize("prec", ["tf32", "3xtf32"])
def test_gemm_split_k(prec):
    from idrk import kernels as K
    M, N, Kc = 512, 512, 5000
    A = _mk((Kc, M), 3)
    B = _mk((Kc, N), 4)
    p = K._PRECISION[prec]
    Au, A_lo = K.split_tf32(A)
    Bu, B_lo = K.split_tf32(B)
    C = torch.zeros(M, N, device=DEV)
    K.gemm(K.GEMM_TN, Au, Bu, M, N, Kc, precision=p, A_lo=A_lo, B_lo=B_lo, C=C, split_k=8)
    ref = A.double().t() @ B.double()
    mag = A.double().abs().t() @ B.double().abs()
    assert ((C.double() - ref).abs() / mag).max().item() < TOL[prec]


def test_gemm_softplus_epilogue_and_split_outputs():
    from idrk import kernels as K
    M, N, Kc = 700, 445, 67
    A, W, b = _mk((M, Kc), 5) * 0.1, _mk((N, Kc), 6) * 0.1, _mk((N,), 7) * 0.01
    Au, A_lo = K.split_tf32(A)
    Wu, W_lo = K.split_tf32(W)
    H, Hh, Hl, S = (K.empty_padded(M, N, DEV) for _ in range(4))
    K.gemm(K.GEMM_NT, Au, Wu, M, N, Kc, precision=K.PREC_3XTF32, A_lo=A_lo, B_lo=W_lo, C=H, C_hi=Hh, C_lo=Hl, S=S,
           bias=b, mode=K.EPI_SOFTPLUS, act=100.0, scale=0.5)
    z = (A.double() @ W.double().t() + b.double())
    ref = torch.nn.functional.softplus(z, beta=100) * 0.5
    assert torch.allclose(H.double(), ref, atol=1e-6, rtol=1e-5)
    assert torch.allclose(S.double(), torch.sigmoid(100 * z), atol=2e-5)
    assert torch.allclose((Hh + Hl).double(), ref, atol=1e-6, rtol=1e-5)


def test_gemm_m_count_masks_rows():
    from idrk import kernels as K
    M, N, Kc = 1000, 128, 64
    A, W = _mk((M, Kc), 8), _mk((N, Kc), 9)
    C = K.empty_padded(M, N, DEV)
    C.fill_(-7.0)
    cnt = torch.tensor([300], device=DEV, dtype=torch.int32)
    K.gemm(K.GEMM_NT, A, W, M, N, Kc, precision=K.PREC_TF32, C=C, m_count=cnt)
    assert (C[300:] == -7.0).all()
    assert torch.allclose(C[:300], A[:300] @ W.t(), atol=0.2, rtol=1e-2)


def test_weight_norm_and_helpers():
    from idrk import kernels as K
    v, g = _mk((445, 512), 10), _mk((445, 1), 11).abs() + 0.1
    out = K.weight_norm_fwd(g, v, True, True)
    ref = torch._weight_norm(v, g, 0)
    assert torch.allclose(out["W"], ref, atol=1e-6, rtol=1e-5)
    assert torch.allclose(out["Wt"], ref.t(), atol=1e-6, rtol=1e-5)
    assert torch.allclose(out["W_hi"] + out["W_lo"], ref, atol=1e-6, rtol=1e-5)
    dW = _mk((445, 512), 12)
    vr, gr = v.clone().requires_grad_(True), g.clone().requires_grad_(True)
    (torch._weight_norm(vr, gr, 0) * dW).sum().backward()
    dg, dv = K.weight_norm_bwd(g, v, dW)
    assert torch.allclose(dg, gr.grad, atol=1e-4, rtol=1e-4)
    assert torch.allclose(dv, vr.grad, atol=1e-5, rtol=1e-4)
    x = _mk((3000, 257), 13)
    assert torch.allclose(K.colsum(x), x.sum(0), atol=1e-3, rtol=1e-4)
    s = _mk((5000,), 14) * 3
    sq, d = K.sdf_squash(s, 0.9001, True)
    beta = 0.9001
    rho = (1 / beta) * (0.5 + 0.5 * s.sign() * torch.expm1(-s.abs() / beta))
    assert torch.allclose(sq, torch.tanh(s / (2 + rho)), atol=1e-6)
    assert torch.allclose(d, (1 - sq * sq) / (2 + rho), atol=1e-6)


@pytest.mark.parametrize("prec", ["tf32", "3xtf32"])
@pytest.mark.parametrize("shape", [(16384, 512, 512), (20000, 445, 67), (33000, 257, 512), (17000, 512, 40)])
def test_gemm_cta_pair_kernel(prec, shape):
    """Large NT launches take the cta_group::2 kernel (256 x 256 tiles per CTA pair)."""
    M, N, Kc = shape
    C, ref, mag = _run(0, prec, M, N, Kc)
    err = ((C.double() - ref).abs() / (mag + 1e-30)).max().item()
    assert err < TOL[prec], err


def test_gemm_cta_pair_epilogue_and_count():
    from idrk import kernels as K
    M, N, Kc = 16384, 512, 512
    A, W, b = _mk((M, Kc), 15) * 0.05, _mk((N, Kc), 16) * 0.05, _mk((N,), 17) * 0.01
    Au, A_lo = K.split_tf32(A)
    Wu, W_lo = K.split_tf32(W)
    Hh, Hl = K.empty_padded(M, N, DEV), K.empty_padded(M, N, DEV)
    Hh.fill_(-3.0)
    Hl.zero_()
    cnt = torch.tensor([9001], device=DEV, dtype=torch.int32)
    K.gemm(K.GEMM_NT, Au, Wu, M, N, Kc, precision=K.PREC_3XTF32, A_lo=A_lo, B_lo=W_lo, C_hi=Hh, C_lo=Hl, bias=b,
           mode=K.EPI_SOFTPLUS, act=100.0, scale=0.70710678, m_count=cnt)
    z = (A[:9001].double() @ W.double().t() + b.double())
    ref = torch.nn.functional.softplus(z, beta=100) * 0.70710678
    assert torch.allclose((Hh + Hl)[:9001].double(), ref, atol=1e-6, rtol=1e-5)
    assert (Hh[9216:] == -3.0).all()            # rows of tiles entirely beyond the count are untouched


@pytest.mark.parametrize("shape", [(128, 128, 64), (300, 445, 67), (4096, 512, 512), (33000, 512, 27), (20000, 512, 512)])
def test_gemm_fp16_pair(shape):
    """fp16-pair operands (x ~= h + l 2^-11): same accuracy class as 3xTF32 (rel 2e-5 of sum|a||b|)."""
    from idrk import kernels as K
    M, N, Kc = shape
    A, W, b = _mk((M, Kc), 21) * 0.3, _mk((N, Kc), 22) * 0.1, _mk((N,), 23) * 0.01
    Ah, Al = K.split_f16(A)
    Wh, Wl = K.split_f16(W)
    C = K.empty_padded(M, N, DEV)
    Ch, Cl = K.empty_half(M, N, DEV), K.empty_half(M, N, DEV)
    K.gemm_f16s(Ah, Al, Wh, Wl, M, N, Kc, C=C, C_h=Ch, C_l=Cl, bias=b)
    ref = A.double() @ W.double().t() + b.double()
    mag = A.double().abs() @ W.double().abs().t() + 1e-30
    assert ((C.double() - ref).abs() / mag).max().item() < 2e-5
    rec = Ch.double() + Cl.double() / 2048.0
    assert ((rec - ref).abs() / mag).max().item() < 3e-5
    cnt = torch.tensor([min(M, 77)], device=DEV, dtype=torch.int32)
    C2 = K.empty_padded(M, N, DEV)
    C2.fill_(5.0)
    K.gemm_f16s(Ah, Al, Wh, Wl, M, N, Kc, C=C2, bias=b, mode=K.EPI_SOFTPLUS, act=100.0, scale=0.5, m_count=cnt)
    k = int(cnt.item())
    assert torch.allclose(C2[:k].double(), torch.nn.functional.softplus(ref[:k], beta=100) * 0.5, atol=2e-6, rtol=2e-5)
    assert (C2[((k + 127) // 128) * 128:] == 5.0).all()


@pytest.mark.parametrize("shape", [(4096, 512, 512), (33000, 512, 512), (300, 445, 67), (129, 96, 64)])
def test_gemm_fp16_pair_fused_row_dot(shape):
    """Fused SDF head: softplus tile times a weight row, one fp32 partial per 32-column group, nothing else stored.
    Partials vs fp64 (abs 2e-5 of sum|act||w| per group), the finished head (sum of partials + bias, Laplace squash)
    vs the two-launch path (activation stored, idrk_sdf_head over K columns), bit-identical run to run, rows beyond
    the device count untouched."""
    from idrk import kernels as K
    M, N, Kc = shape
    A, W, b = _mk((M, Kc), 41) * 0.3, _mk((N, Kc), 42) * 0.1, _mk((N,), 43) * 0.01
    w = _mk((N,), 44)
    b0 = _mk((1,), 45)
    Ah, Al = K.split_f16(A)
    Wh, Wl = K.split_f16(W)
    G = (N + 31) // 32
    cnt_v = min(M, 3001)
    cnt = torch.tensor([cnt_v], device=DEV, dtype=torch.int32)
    part = torch.full((M, K.pad4(G)), 9.0, device=DEV)
    K.gemm_f16s(Ah, Al, Wh, Wl, M, N, Kc, bias=b, mode=K.EPI_SOFTPLUS, act=100.0, scale=0.70710678, m_count=cnt,
                dot_w=w, dot_out=part)
    act = torch.nn.functional.softplus(A.double() @ W.double().t() + b.double(), beta=100) * 0.70710678
    prod = act * w.double()
    padc = G * 32 - N
    prod_p = torch.nn.functional.pad(prod, (0, padc))
    ref = prod_p.reshape(M, G, 32).sum(-1)
    mag = prod_p.abs().reshape(M, G, 32).sum(-1) + 1e-30
    got = part[:cnt_v, :G].double()
    assert ((got - ref[:cnt_v]).abs() / mag[:cnt_v]).max().item() < 2e-5
    assert (part[((cnt_v + 127) // 128) * 128:] == 9.0).all()
    part2 = torch.full_like(part, 9.0)
    K.gemm_f16s(Ah, Al, Wh, Wl, M, N, Kc, bias=b, mode=K.EPI_SOFTPLUS, act=100.0, scale=0.70710678, m_count=cnt,
                dot_w=w, dot_out=part2)
    assert torch.equal(part[:cnt_v, :G], part2[:cnt_v, :G])
    # finished head vs the unfused pair of launches
    out_f = torch.zeros(M, device=DEV)
    K.sdf_head(part, torch.ones(G, device=DEV), b0, 0.9001, out_f, M, cnt)
    H = K.empty_padded(M, N, DEV)
    K.gemm_f16s(Ah, Al, Wh, Wl, M, N, Kc, C=H, bias=b, mode=K.EPI_SOFTPLUS, act=100.0, scale=0.70710678, m_count=cnt)
    out_u = torch.zeros(M, device=DEV)
    K.sdf_head(H, w, b0, 0.9001, out_u, M, cnt)
    assert (out_f[:cnt_v] - out_u[:cnt_v]).abs().max().item() <= 2e-6 * max(1.0, mag.sum(-1).max().item())
    # argument errors: the activation cannot be stored at the same time; partial rows too short
    from idrk._lib import IdrkError
    with pytest.raises(IdrkError):
        K.gemm_f16s(Ah, Al, Wh, Wl, M, N, Kc, C=H, bias=b, mode=K.EPI_SOFTPLUS, act=100.0, dot_w=w, dot_out=part)
    if G > 1:
        with pytest.raises(IdrkError):
            K.gemm_f16s(Ah, Al, Wh, Wl, M, N, Kc, bias=b, mode=K.EPI_SOFTPLUS, act=100.0, dot_w=w,
                        dot_out=torch.zeros(M, G - 1, device=DEV))


def test_split_f16_two_destinations():
    """One launch writes the fp16 pair of x and, into a column slot of another buffer, the pair of scale2 * x."""
    from idrk import kernels as K
    x = _mk((777, 27), 31) * 3.0
    xp = K.operand(x)
    h, l = K.empty_half(777, 27, DEV), K.empty_half(777, 27, DEV)
    big_h = torch.full((777, 512), 7.0, device=DEV, dtype=torch.float16)
    big_l = torch.full((777, 512), 7.0, device=DEV, dtype=torch.float16)
    cnt = torch.tensor([700], device=DEV, dtype=torch.int32)
    K.split_f16_into(xp, 777, 27, 1.0, h, l, K.pad8(27), K.pad8(27) - 27, cnt,
                     second=(big_h[:, 485:], big_l[:, 485:], 512, 0, 0.5))
    rec = h[:700, :27].double() + l[:700, :27].double() / 2048.0
    assert ((rec - x[:700].double()).abs() / x[:700].double().abs().clamp_min(1e-6)).max().item() < 1e-6
    rec2 = big_h[:700, 485:512].double() + big_l[:700, 485:512].double() / 2048.0
    assert ((rec2 - 0.5 * x[:700].double()).abs() / x[:700].double().abs().clamp_min(1e-6)).max().item() < 1e-6
    assert (big_h[:, :485] == 7.0).all() and (big_h[700:] == 7.0).all()      # nothing outside the slot / the count
    assert (h[:700, 27:] == 0).all()                                             # pad columns zero-filled


# ---------------------------------------------------------------------------------------------
# 16-bit-pair contraction of the differentiable path (csrc/gemm_p16.cu): fp16 pairs ~22 bits, bf16 pairs ~17 bits
# ---------------------------------------------------------------------------------------------
P16_TOL = {(0, 0): 2e-6, (1, 1): 2e-5}      # relative to sum |a||b| per output


def _run_p16(layout, M, N, Kc, fa, fb, scale_a=1.0, scale_b=1.0, **kw):
    from idrk import kernels as K
    A = (_mk((M, Kc), 1) if layout != K.GEMM_TN else _mk((Kc, M), 1)) * scale_a
    B = (_mk((N, Kc), 2) if layout == K.GEMM_NT else _mk((Kc, N), 2)) * scale_b
    C = K.empty_padded(M, N, DEV)
    C.fill_(float("nan"))
    K.gemm_p16(layout, K.split_p16(A, fa), K.split_p16(B, fb), M, N, Kc, C=C, **kw)
    A64 = A.double() if layout != K.GEMM_TN else A.double().t()
    B64 = B.double() if layout == K.GEMM_NT else B.double().t()
    return C, A64 @ B64.t(), A64.abs() @ B64.abs().t()


@pytest.mark.parametrize("fmts", [(0, 0), (1, 1)])
@pytest.mark.parametrize("layout", [0, 1, 2])
@pytest.mark.parametrize("shape", [(128, 128, 64), (300, 445, 67), (2048, 512, 512), (4096, 512, 512), (77, 3, 512),
                                   (1000, 257, 40), (4096, 256, 39)])
def test_gemm_p16_plain(layout, fmts, shape):
    M, N, Kc = shape
    C, ref, mag = _run_p16(layout, M, N, Kc, *fmts)
    err = ((C.double() - ref).abs() / (mag + 1e-30)).max().item()
    assert err < P16_TOL[fmts], err


def test_gemm_p16_bf16_pair_keeps_the_fp32_range():
    """Cotangent-sized operands (1e-9 ... 1e-12) that fp16 pairs would flush: bf16 pairs keep 17 bits at any magnitude."""
    C, ref, mag = _run_p16(1, 1024, 512, 512, 1, 1, scale_a=1e-9)
    assert ((C.double() - ref).abs() / mag).max().item() < 2e-5
    C, ref, mag = _run_p16(2, 512, 512, 2048, 1, 1, scale_a=1e-12, scale_b=1e6)
    assert ((C.double() - ref).abs() / mag).max().item() < 2e-5


@pytest.mark.parametrize("fmts", [(0, 0), (1, 1)])
def test_gemm_p16_split_k(fmts):
    from idrk import kernels as K
    M, N, Kc = 512, 512, 5000
    A, B = _mk((Kc, M), 3), _mk((Kc, N), 4)
    C = torch.zeros(M, N, device=DEV)
    K.gemm_p16(K.GEMM_TN, K.split_p16(A, fmts[0]), K.split_p16(B, fmts[1]), M, N, Kc, C=C, split_k=8)
    ref = A.double().t() @ B.double()
    mag = A.double().abs().t() @ B.double().abs()
    assert ((C.double() - ref).abs() / mag).max().item() < P16_TOL[fmts]


def test_gemm_p16_rejects_mixed_formats():
    from idrk import kernels as K
    from idrk._lib import IdrkError
    A, B = _mk((128, 64), 1), _mk((128, 64), 2)
    with pytest.raises(IdrkError):
        K.gemm_p16(K.GEMM_NT, K.split_p16(A, 0), K.split_p16(B, 1), 128, 128, 64, C=K.empty_padded(128, 128, DEV))


@pytest.mark.parametrize("c_fmt", [0, 1])
@pytest.mark.parametrize("shape", [(700, 445, 67), (4096, 512, 512), (2048, 512, 39)])
def test_gemm_p16_softplus_epilogue_and_pair_output(c_fmt, shape):
    from idrk import kernels as K
    M, N, Kc = shape
    A, W, b = _mk((M, Kc), 5) * 0.1, _mk((N, Kc), 6) * 0.1, _mk((N,), 7) * 0.01
    H, S = K.empty_padded(M, N, DEV), K.empty_padded(M, N, DEV)
    Ch, Cl = K.empty_pair16(M, N, DEV, c_fmt)
    K.gemm_p16(K.GEMM_NT, K.split_p16(A, 1), K.split_p16(W, 1), M, N, Kc, C=H, C_pair=(Ch, Cl, c_fmt), S=S, bias=b,
               mode=K.EPI_SOFTPLUS, act=100.0, scale=0.5)
    z = (A.double() @ W.double().t() + b.double())
    ref = torch.nn.functional.softplus(z, beta=100) * 0.5
    assert torch.allclose(H.double(), ref, atol=5e-6, rtol=2e-5)
    assert torch.allclose(S.double(), torch.sigmoid(100 * z), atol=5e-4)
    pair = Ch.double() + Cl.double() / 2048.0
    assert torch.allclose(pair, H.double(), atol=1e-9, rtol=3e-7 if c_fmt == 0 else 1e-5)


@pytest.mark.parametrize("mode", ["relu", "sine", "tanh", "none"])
def test_gemm_p16_other_epilogues(mode):
    from idrk import kernels as K
    M, N, Kc = 1000, 256, 120
    A, W, b = _mk((M, Kc), 15) * 0.2, _mk((N, Kc), 16) * 0.2, _mk((N,), 17) * 0.1
    H, S = K.empty_padded(M, N, DEV), K.empty_padded(M, N, DEV)
    em = {"relu": K.EPI_RELU, "sine": K.EPI_SINE, "tanh": K.EPI_TANH, "none": K.EPI_NONE}[mode]
    K.gemm_p16(K.GEMM_NT, K.split_p16(A, 0), K.split_p16(W, 0), M, N, Kc, C=H, S=S, bias=b, mode=em, act=30.0, scale=1.0)
    z = (A.double() @ W.double().t() + b.double())
    ref, dref = {"relu": (z.clamp_min(0), (z > 0).double()), "sine": (torch.sin(30 * z), 30 * torch.cos(30 * z)),
                 "tanh": (torch.tanh(z), 1 - torch.tanh(z) ** 2), "none": (z, torch.ones_like(z))}[mode]
    tol = 2e-4 if mode == "sine" else 5e-6
    assert (H.double() - ref).abs().max().item() < tol
    if mode != "relu":
        assert (S.double() - dref).abs().max().item() < (1e-2 if mode == "sine" else 1e-5)
    else:
        far = z.abs() > 1e-5
        assert torch.equal(S.double()[far], dref[far])


def test_p16_helpers_match_the_fp32_kernels():
    """weight_norm / act_bwd with a 16-bit pair output: fp32 results identical to the tf32-pair entry points, pair == value."""
    from idrk import kernels as K
    v, g = _mk((445, 512), 10), _mk((445, 1), 11).abs() + 0.1
    W, (h, l, fmt) = K.weight_norm_fwd_p16(g, v, 0)
    ref = K.weight_norm_fwd(g, v, False, False)["W"]
    assert torch.equal(W, ref)
    assert torch.allclose(h.double() + l.double() / 2048, ref.double(), atol=1e-9, rtol=3e-7)
    dH, S = _mk((3000, 445), 20) * 1e-6, torch.rand(3000, 445, device=DEV)
    dS, H = _mk((3000, 445), 21) * 1e-7, _mk((3000, 445), 22)
    for mode in (K.EPI_SOFTPLUS, K.EPI_TANH, K.EPI_RELU):
        a = K.act_bwd(dH, dS if mode != K.EPI_RELU else None, S, H, mode, 100.0, 0.5, False)[0]
        b, (bh, bl, _) = K.act_bwd_p16(dH, dS if mode != K.EPI_RELU else None, S, H, mode, 100.0, 0.5, 1)
        assert torch.equal(a, b)
        assert torch.allclose(bh.double() + bl.double() / 2048, a.double(), atol=0, rtol=1e-5)
