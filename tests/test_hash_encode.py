"""GPU parity of K1/K2 (hash-grid encode fwd/bwd, posenc) against the oracle and the golden vectors.
Tolerances: hash indices and hash features bit-exact; Fourier/posenc columns abs <= 4e-6 (accurate
sinf/cosf vs the host libm); table gradients rel <= 1e-5 of max-abs (fp32 atomics reorder sums)."""
import numpy as np
import pytest
import torch

from conftest import sd_from
from oracle import idr_oracle as O
from tests_support import load_sd_into

pytestmark = pytest.mark.gpu
DEV = "cuda"


def T(a):
    return torch.from_numpy(np.array(a))


def make_grid(L, F, log2T, base, desired, mode="reference", seed=0, std=0.5):
    from idrk.model.embeddings.hashGridEmbedding import MultiResHashGridMLP
    gen = torch.Generator().manual_seed(seed)
    sd = O.make_hashgrid_sd("", L, F, log2T, base, desired, gen, table_std=std)
    m = MultiResHashGridMLP(True, 3, L, F, log2T, base, desired, frac_mode=mode)
    load_sd_into(m, sd)
    return m.to(DEV), sd


@pytest.mark.parametrize("tag", ["e1", "e2"])
def test_golden_embedding_and_grads(golden, tag):
    from idrk.model.embeddings.hashGridEmbedding import MultiResHashGridMLP
    g = golden("hashgrid")
    L, F, log2T, base, desired = [int(v) for v in g["emb_args_" + tag]]
    m = MultiResHashGridMLP(True, 3, L, F, log2T, base, desired)
    load_sd_into(m, sd_from(g, "sd_%s/" % tag))
    m = m.to(DEV)
    x = T(g["x"]).to(DEV)
    y = m(x)
    ref = T(g["emb_" + tag])
    assert y.shape == ref.shape
    assert torch.equal(y[:, 3 + 2 * L:].cpu(), ref[:, 3 + 2 * L:])
    assert torch.allclose(y[:, :3 + 2 * L].cpu(), ref[:, :3 + 2 * L], atol=4e-6, rtol=0)
    (y * T(g["w_" + tag]).to(DEV)).sum().backward()
    for l, lvl in enumerate(m.levels):
        r = T(g["grad_%s/%d" % (tag, l)])
        assert torch.allclose(lvl.embedding.weight.grad.cpu(), r, atol=1e-5 * max(r.abs().max().item(), 1e-9), rtol=1e-5)


def test_golden_corner_indices_bit_exact(golden):
    from idrk import kernels as K
    g = golden("hashgrid")
    x = T(g["x"]).to(DEV)
    for tag in ("a", "b", "c"):
        res, rows = [int(v) for v in g["meta_" + tag]]
        spec = K.HashGridSpec([res], [rows], 2, 0, 0)
        table = torch.zeros(rows, 2, device=DEV)
        _, idx = K.hash_encode_fwd(spec, x, (table,), None, want_idx=True)
        got = idx[:, 0, :].cpu().numpy().astype(np.int64) & 0xFFFFFFFF
        assert np.array_equal(got, g["idx_" + tag])


@pytest.mark.parametrize("n", [0, 1, 31, 257, 8192, 100003])
def test_reference_mode_vs_oracle_sizes(n):
    m, sd = make_grid(16, 2, 19, 16, 2048)
    x = torch.rand(n, 3, generator=torch.Generator().manual_seed(n)) * 2 - 1
    y = m(x.to(DEV)).cpu()
    assert y.shape == (n, 67)
    if n == 0:
        return
    ref = O.hashgrid_embed(x, sd, "", 16, 16, 2048)
    assert torch.equal(y[:, 35:], ref[:, 35:])
    assert torch.allclose(y[:, :35], ref[:, :35], atol=4e-6, rtol=0)


@pytest.mark.parametrize("F", [1, 2, 4, 8])
@pytest.mark.parametrize("mode", ["reference", "trilinear"])
def test_feature_widths_and_modes(F, mode):
    m, sd = make_grid(5, F, 9, 8, 96, mode=mode, seed=3)
    x = torch.rand(3001, 3, generator=torch.Generator().manual_seed(1)) * 1.6 - 0.3
    y = m(x.to(DEV)).cpu()
    ref = O.hashgrid_embed(x, sd, "", 5, 8, 96, mode)
    pre = 3 + 2 * 5
    if mode == "reference":
        assert torch.equal(y[:, pre:], ref[:, pre:])
    else:
        assert torch.allclose(y[:, pre:], ref[:, pre:], atol=2e-6, rtol=1e-5)


@pytest.mark.parametrize("mode", ["reference", "trilinear"])
def test_backward_tables_and_dx(mode):
    L, F = 8, 2
    m, sd = make_grid(L, F, 10, 8, 256, mode=mode, seed=5)
    gen = torch.Generator().manual_seed(2)
    x = torch.rand(5000, 3, generator=gen) * 2 - 1
    w = torch.randn(5000, 3 + 2 * L + L * F, generator=gen)
    xd = x.to(DEV).requires_grad_(True)
    (m(xd) * w.to(DEV)).sum().backward()
    for v in sd.values():
        v.requires_grad_(v.dim() == 2 and v.shape[0] != 3)
    xo = x.clone().requires_grad_(True)
    (O.hashgrid_embed(xo, sd, "", L, 8, 256, mode) * w).sum().backward()
    for l, lvl in enumerate(m.levels):
        r = sd["levels.%d.embedding.weight" % l].grad
        got = lvl.embedding.weight.grad.cpu()
        assert torch.allclose(got, r, atol=1e-5 * r.abs().max().item(), rtol=1e-4), l
    if mode == "reference":
        assert torch.allclose(xd.grad.cpu(), xo.grad, atol=2e-5 * xo.grad.abs().max().item(), rtol=1e-4)


@pytest.mark.parametrize("mode", ["reference", "trilinear"])
def test_aligned_store_paths_l16(mode):
    """L = C = 16, F = 2: a lane's level is fixed and the level columns start on an odd column, so the forward
    stores / backward loads take the shuffle-realigned float2 path.  Forward, table grads and dL/dx vs the oracle,
    at a ragged size (last warp tile partially filled)."""
    L, F, n = 16, 2, 4099
    m, sd = make_grid(L, F, 12, 16, 512, mode=mode, seed=11)
    gen = torch.Generator().manual_seed(6)
    x = torch.rand(n, 3, generator=gen) * 2 - 1
    w = torch.randn(n, 3 + 2 * L + L * F, generator=gen)
    xd = x.to(DEV).requires_grad_(True)
    y = m(xd)
    (y * w.to(DEV)).sum().backward()
    for v in sd.values():
        v.requires_grad_(v.dim() == 2 and v.shape[0] != 3)
    xo = x.clone().requires_grad_(True)
    ref = O.hashgrid_embed(xo, sd, "", L, 16, 512, mode)
    (ref * w).sum().backward()
    pre = 3 + 2 * L
    assert torch.allclose(y[:, :pre].detach().cpu(), ref[:, :pre].detach(), atol=4e-6, rtol=0)
    if mode == "reference":
        assert torch.equal(y[:, pre:].detach().cpu(), ref[:, pre:].detach())
    else:
        assert torch.allclose(y[:, pre:].detach().cpu(), ref[:, pre:].detach(), atol=2e-6, rtol=1e-5)
    for l, lvl in enumerate(m.levels):
        r = sd["levels.%d.embedding.weight" % l].grad
        assert torch.allclose(lvl.embedding.weight.grad.cpu(), r, atol=1e-5 * r.abs().max().item(), rtol=1e-4), l
    if mode == "reference":     # trilinear dL/dx is discontinuous at cell faces; covered by test_feature_widths_and_modes
        assert torch.allclose(xd.grad.cpu(), xo.grad, atol=2e-5 * xo.grad.abs().max().item(), rtol=1e-4)
    else:
        bad = ~torch.isclose(xd.grad.cpu(), xo.grad, atol=1e-3 * xo.grad.abs().max().item(), rtol=1e-3)
        assert bad.sum().item() <= 3


def test_trilinear_paired_corners_mixed_levels():
    """8-corner mode, F = 2: x-neighbour corners of an even cell coordinate are fetched / reduced as ONE 16-byte
    access when the level's row count is even and its (gradient) table is 16-byte aligned.  Here L = 8 levels mix odd
    row counts (15^3, 33^3) with even ones, and a second run puts every table and gradient table on an 8-byte-only
    boundary (views into a flat buffer, as in the trainer's bucket), so paired and unpaired lanes share warps and the
    fall-back is taken.  Forward must equal the oracle to rounding and be bit-identical between the two placements."""
    from idrk import kernels as K
    L, F, n = 8, 2, 4096 + 77
    m, sd = make_grid(L, F, 16, 15, 255, mode="trilinear", seed=21)
    rows = [lvl.embedding.weight.shape[0] for lvl in m.levels]
    assert any(r % 2 for r in rows) and any(r % 2 == 0 for r in rows)
    gen = torch.Generator().manual_seed(8)
    x = torch.rand(n, 3, generator=gen) * 2 - 1
    w = torch.randn(n, K.pad4(3 + 2 * L + L * F), generator=gen)
    w[:, 3 + 2 * L + L * F:] = 0
    for v in sd.values():
        v.requires_grad_(v.dim() == 2 and v.shape[0] != 3)
    ref = O.hashgrid_embed(x, sd, "", L, 15, 255, "trilinear")
    (ref * w[:, :ref.shape[1]]).sum().backward()
    spec, B = m.spec(), m.freq_encoding.B
    xd, wd = x.to(DEV), w.to(DEV)
    outs = []
    for shift in (0, 2):                    # floats: 0 -> 16-byte aligned tables, 2 -> 8-byte-only
        tabs, grads = [], []
        for lvl in m.levels:
            t = lvl.embedding.weight.detach()
            flat = torch.zeros(t.numel() + 4, device=DEV)
            gflat = torch.zeros(t.numel() + 4, device=DEV)
            tv = flat[shift:shift + t.numel()].view_as(t)
            tv.copy_(t)
            tabs.append(tv)
            grads.append(gflat[shift:shift + t.numel()].view_as(t))
            assert tv.data_ptr() % 16 == 4 * shift
        y = K.hash_encode_fwd(spec, xd, tuple(tabs), B)
        K.hash_encode_bwd(spec, xd, tuple(tabs), B, wd, grads, False)
        pre = 3 + 2 * L
        assert torch.allclose(y[:, pre:ref.shape[1]].cpu(), ref[:, pre:].detach(), atol=2e-6, rtol=1e-5)
        for l in range(L):
            r = sd["levels.%d.embedding.weight" % l].grad
            assert torch.allclose(grads[l].cpu(), r, atol=1e-5 * r.abs().max().item(), rtol=1e-4), (shift, l)
        outs.append(y)
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("mode", ["reference", "trilinear", "ngp"])
def test_encode_to_fp16_pair_equals_encode_then_split(mode):
    """K1p (one launch: encode -> fp16 pair, + the scaled second copy) == K1 followed by split_f16, bit for bit, for
    the first *m_count rows; rows beyond the count and columns outside the slot untouched; pad columns zero."""
    from idrk import kernels as K
    n, cnt_v = 3000, 2777
    gen = torch.Generator().manual_seed(4)
    if mode == "ngp":
        from idrk.model.embeddings.tcnn_src.hashGridEncoderTcnn import NgpGrid
        m = NgpGrid(6, 2, 12, 16, 2.0)
        with torch.no_grad():
            m.params.copy_(torch.randn(m.params.shape, generator=gen))
        m = m.to(DEV)
        spec, tables, B = m.spec(), m.tables(), None
        x = torch.rand(n, 3, generator=gen).to(DEV)
    else:
        m, _ = make_grid(6, 2, 5, 64, 512, mode=mode, seed=9)
        spec, tables, B = m.spec(), m.tables(), m.freq_encoding.B
        x = (torch.rand(n, 3, generator=gen) * 2 - 1).to(DEV)
    E = spec.width
    cnt = torch.tensor([cnt_v], device=DEV, dtype=torch.int32)
    emb = K.hash_encode_fwd(spec, x, tables, B)
    ld, big = K.pad8(E), 512
    h0, l0 = torch.full((n, ld), 7.0, device=DEV, dtype=torch.float16), torch.full((n, ld), 7.0, device=DEV, dtype=torch.float16)
    bh0, bl0 = torch.full((n, big), 7.0, device=DEV, dtype=torch.float16), torch.full((n, big), 7.0, device=DEV, dtype=torch.float16)
    h1, l1, bh1, bl1 = h0.clone(), l0.clone(), bh0.clone(), bl0.clone()
    off = big - K.pad8(E)
    K.split_f16_into(emb, n, E, 1.0, h0, l0, ld, ld - E, cnt, second=(bh0[:, off:], bl0[:, off:], big, K.pad8(E) - E, 0.70710678))
    K.hash_encode_f16pair(spec, x, tables, B, n, h1, l1, ld, ld - E, cnt, second=(bh1[:, off:], bl1[:, off:], big, K.pad8(E) - E, 0.70710678))
    for a, b in ((h0, h1), (l0, l1), (bh0, bh1), (bl0, bl1)):
        assert torch.equal(a, b)
    assert (h1[cnt_v:] == 7.0).all() and (bh1[:, :off] == 7.0).all() and (h1[:cnt_v, E:] == 0).all()
    rec = h1[:cnt_v, :E].double() + l1[:cnt_v, :E].double() / 2048.0
    assert (rec - emb[:cnt_v, :E].double()).abs().max().item() <= 1e-6 * max(1.0, emb.abs().max().item())


def test_unpadded_rows_and_huge_coordinates():
    """ld_out == width (odd: no vector path, no pad column) and coordinates whose scaled value leaves the int32
    range (the .long() emulation has to take the 64-bit conversion): indices stay bit-exact."""
    from idrk import kernels as K
    L, F = 16, 2
    m, sd = make_grid(L, F, 12, 16, 512, seed=13)
    x = torch.rand(1000, 3, generator=torch.Generator().manual_seed(8)) * 2 - 1
    x[::7] *= 3.0e7
    x[5, 1] = -2.5e9
    spec, tables, B = m.spec(), tuple(t.detach() for t in m.tables()), m.freq_encoding.B
    out = torch.full((1000, spec.width), float("nan"), device=DEV)
    K.hash_encode_fwd(spec, x.to(DEV), tables, B, out=out)
    ref = O.hashgrid_embed(x, sd, "", L, 16, 512)
    assert torch.equal(out[:, 35:].cpu(), ref[:, 35:])
    assert torch.equal(out[:, :3].cpu(), x)


def test_tiny_tables_shared_accumulation():
    """Reference configs use T = 32 rows per level: every point collides (CTA-local accumulators)."""
    L, F = 6, 2
    m, sd = make_grid(L, F, 5, 64, 512, seed=7)
    gen = torch.Generator().manual_seed(4)
    x = torch.rand(20000, 3, generator=gen) * 2 - 1
    w = torch.randn(20000, 3 + 2 * L + L * F, generator=gen)
    (m(x.to(DEV)) * w.to(DEV)).sum().backward()
    for v in sd.values():
        v.requires_grad_(v.dim() == 2 and v.shape[0] != 3)
    (O.hashgrid_embed(x, sd, "", L, 64, 512) * w).sum().backward()
    for l, lvl in enumerate(m.levels):
        r = sd["levels.%d.embedding.weight" % l].grad
        assert torch.allclose(lvl.embedding.weight.grad.cpu(), r, atol=2e-5 * r.abs().max().item(), rtol=1e-4)


def test_linearity_and_checksum_at_full_size():
    """Size-independent properties at a BASELINE-sized microbench case (4M points, 2^19 table):
    gradient scatter conserves mass (sum of table grads == sum of upstream grads) and the encode is
    a pure gather (every output value is one of the table's values)."""
    L, F = 16, 2
    m, _ = make_grid(L, F, 19, 16, 2048, seed=9)
    n = 1 << 22
    x = torch.rand(n, 3, device=DEV)
    y = m(x)
    dy = torch.randn_like(y)
    y.backward(dy)
    for l, lvl in enumerate(m.levels):
        got = lvl.embedding.weight.grad.double().sum().item()
        want = dy[:, 35 + 2 * l: 37 + 2 * l].double().sum().item()
        assert abs(got - want) <= 1e-3 * max(1.0, abs(want)) + 0.5, (l, got, want)
    lvl0 = m.levels[0].embedding.weight
    vals = torch.unique(lvl0.detach()[:, 0])
    assert torch.isin(y[:4096, 35], vals).all()


def test_posenc_and_fourier(golden):
    from idrk.model.embeddings.frequency_enc import PositionalEncoding, get_embedder, FourierFeature
    g = golden("encoders")
    x = T(g["x"]).to(DEV)
    pe = PositionalEncoding(include_input=True, input_dims=3, max_freq_log2=5, num_freqs=6, log_sampling=True,
                            periodic_fns=[torch.sin, torch.cos])
    assert torch.allclose(pe(x).cpu(), T(g["posenc_6_5"]), atol=4e-6)
    fn, od = get_embedder(4)
    assert od == int(g["view_nerfpos4_outdim"][0])
    assert torch.allclose(fn(x).cpu(), T(g["view_nerfpos4"]), atol=4e-6)
    xr = x.clone().requires_grad_(True)
    w = torch.randn(x.shape[0], 42, device=DEV)
    (pe(xr) * w).sum().backward()
    xo = T(g["x"]).clone().requires_grad_(True)
    (O.positional_encoding(xo, 6, 5, True) * w.cpu()).sum().backward()
    assert torch.allclose(xr.grad.cpu(), xo.grad, atol=1e-4 * xo.grad.abs().max().item())
    ff = FourierFeature(3, 1.0, 3).to(DEV)
    y = ff(x)
    ref = O.fourier_feature(T(g["x"]), ff.B.cpu())
    assert torch.allclose(y.cpu(), ref, atol=4e-6)


@pytest.mark.parametrize("n", [31, 4096 + 17, 200000])
def test_trilinear_backward_run_aggregation(n):
    """8-corner table-gradient backward with in-lane run aggregation (csrc/hash_encode.cu): Z-ordered points (long runs on
    the coarse levels), the same points in random order (no runs) and a batch of identical points (one run per lane) must
    give the same table gradients - rel 1e-5 of max-abs, fp32 sums reorder - and match the oracle's dense gradient."""
    from idrk import kernels as K
    from idrk.model.embeddings.hashGridEmbedding import MultiResHashGridMLP
    from idrk.utils.sorting import morton_order
    L, F, log2T = 16, 2, 14
    gen = torch.Generator().manual_seed(n)
    m = MultiResHashGridMLP(True, 3, L, F, log2T, 16, 2048, frac_mode="trilinear").to(DEV)
    spec, tables, B = m.spec(), tuple(t.detach() for t in m.tables()), m.freq_encoding.B
    x = torch.rand(n, 3, generator=gen).to(DEV)
    dy = torch.randn(n, K.pad4(spec.width), generator=gen).to(DEV)

    def grads(xx, dd, ordered=True):
        gt = [torch.zeros_like(t) for t in tables]
        K.hash_encode_bwd(spec, xx.contiguous(), tables, B, dd.contiguous(), gt, False, ordered=ordered)
        return gt
    g_plain = grads(x, dy, ordered=False)                  # the plain kernel (no aggregation code)
    g_rand = grads(x, dy)                                  # aggregating kernel on unordered input: no runs
    for l, (a, b) in enumerate(zip(g_plain, g_rand)):
        assert (a - b).abs().max().item() <= 1e-5 * max(a.abs().max().item(), 1e-12), l
    perm = morton_order(x, lo=(0.0, 0.0, 0.0), hi=(1.0, 1.0, 1.0))
    g_sort = grads(x[perm], dy[perm])
    for l, (a, b) in enumerate(zip(g_rand, g_sort)):
        assert (a - b).abs().max().item() <= 1e-5 * max(a.abs().max().item(), 1e-12), l
    # all points identical: every lane aggregates its whole tile range into one cell
    x1 = x[:1].expand(n, 3).contiguous()
    g_same = grads(x1, dy)
    s = dy[:, 3 + 2 * L:3 + 2 * L + L * F].double().sum(0)           # per (level, feature) column sums of dL/dy
    for l, gtab in enumerate(g_same):
        tot = gtab.double().sum(0).cpu()                              # corner weights sum to 1: mass is conserved
        ref = s[l * F:(l + 1) * F].cpu()
        assert (tot - ref).abs().max().item() <= 1e-4 * max(ref.abs().max().item(), 1.0), l
        assert (gtab != 0).any(1).sum().item() <= 8, l               # ... and lands in at most 8 rows
    if n <= 5000:                                                     # oracle: dense autograd gradient on the CPU
        sd = {"levels.%d.embedding.weight" % l: t.detach().cpu().clone().requires_grad_(True) for l, t in enumerate(tables)}
        sd["freq_encoding.B"] = B.cpu()
        y = O.hashgrid_embed(x[perm].cpu(), sd, "", L, 16, 2048, "trilinear")
        (y * dy[perm, :spec.width].cpu()).sum().backward()
        for l, gtab in enumerate(g_sort):
            ref = sd["levels.%d.embedding.weight" % l].grad
            assert (gtab.cpu() - ref).abs().max().item() <= 1e-5 * max(ref.abs().max().item(), 1e-12), l


@pytest.mark.parametrize("n,bits", [(1, 8), (31, 10), (4096, 8), (4097, 3), (100003, 8), (1 << 20, 10)])
def test_morton_sort_is_a_stable_z_order_permutation(n, bits):
    """csrc/point_sort.cu vs a host reference: keys = bit-interleaved lattice coordinates, stable ascending sort.
    Integer work: the permutation must be IDENTICAL to numpy's stable argsort of the same keys."""
    from idrk import kernels as K
    gen = torch.Generator().manual_seed(n + bits)
    x = torch.rand(n, 3, generator=gen) * 1.4 - 0.2            # some points outside the box: clamped
    if n > 100:
        x[::7] = x[3]                                          # many equal keys: stability matters
    perm = K.morton_perm(x.to(DEV), lo=(0.0, 0.0, 0.0), hi=(1.0, 1.0, 1.0), bits=bits).cpu().numpy()
    q = np.floor(np.clip(x.numpy().astype(np.float32) * np.float32(2 ** bits), 0, 2 ** bits - 1)).astype(np.uint64)

    def spread(v):
        out = np.zeros_like(v)
        for b in range(10):
            out |= ((v >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b)
        return out
    keys = spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1)) | (spread(q[:, 2]) << np.uint64(2))
    ref = np.argsort(keys, kind="stable")
    assert np.array_equal(np.sort(perm), np.arange(n))
    assert np.array_equal(perm, ref)


@pytest.mark.parametrize("mode", ["reference", "trilinear"])
@pytest.mark.parametrize("n", [33, 70001])
def test_encode_through_a_permutation(mode, n):
    """`perm` only changes the ORDER in which points are processed: the forward output is bit-identical to the plain
    call (rows stay where they were), table gradients equal up to fp32 summation order (rel 1e-5 of max), dL/dx equal."""
    from idrk import kernels as K
    from idrk.model.embeddings.hashGridEmbedding import MultiResHashGridMLP
    L, F, log2T = 16, 2, 12
    gen = torch.Generator().manual_seed(n)
    m = MultiResHashGridMLP(True, 3, L, F, log2T, 16, 2048, frac_mode=mode).to(DEV)
    with torch.no_grad():
        for t in m.tables():
            t.mul_(3000.0)
    spec, tables, B = m.spec(), tuple(t.detach() for t in m.tables()), m.freq_encoding.B
    x = torch.rand(n, 3, generator=gen).to(DEV)
    dy = torch.randn(n, K.pad4(spec.width), generator=gen).to(DEV)
    perm = K.morton_perm(x)
    y0 = K.hash_encode_fwd(spec, x, tables, B)
    y1 = K.hash_encode_fwd(spec, x, tables, B, perm=perm)
    assert torch.equal(y0, y1)
    for want_dx in (False, True):
        g0 = [torch.zeros_like(t) for t in tables]
        g1 = [torch.zeros_like(t) for t in tables]
        d0 = K.hash_encode_bwd(spec, x, tables, B, dy, g0, want_dx)
        d1 = K.hash_encode_bwd(spec, x, tables, B, dy, g1, want_dx, perm=perm)
        for l, (a, b) in enumerate(zip(g0, g1)):
            assert (a - b).abs().max().item() <= 1e-5 * max(a.abs().max().item(), 1e-12), (want_dx, l)
        if want_dx:
            assert (d0 - d1).abs().max().item() <= 1e-5 * max(d0.abs().max().item(), 1e-12)


def test_module_sorts_big_unordered_batches_transparently():
    """MultiResHashGridMLP(frac_mode='trilinear') on >= 2^18 points walks them in Z-order (autograd_ops.AUTO_SORT): outputs
    bit-identical to the unsorted walk, table gradients and dL/dx equal up to fp32 summation order; reference mode never sorts."""
    from idrk import autograd_ops as ops
    n = (1 << 18) + 77
    gen = torch.Generator().manual_seed(1)
    m, _ = make_grid(16, 2, 14, 16, 2048, mode="trilinear", seed=2)
    x = torch.rand(n, 3, generator=gen).to(DEV)
    w = torch.randn(n, m.embeddings_dim, generator=gen).to(DEV)

    def run(enabled):
        ops.AUTO_SORT["enabled"] = enabled
        try:
            xx = x.clone().requires_grad_(True)
            for t in m.tables():
                t.grad = None
            y = m(xx)
            (y * w).sum().backward()
            return y.detach(), xx.grad.clone(), [t.grad.clone() for t in m.tables()]
        finally:
            ops.AUTO_SORT["enabled"] = True
    y1, dx1, g1 = run(True)
    y0, dx0, g0 = run(False)
    assert torch.equal(y0, y1)
    assert (dx0 - dx1).abs().max().item() <= 1e-5 * dx0.abs().max().item()
    for l, (a, b) in enumerate(zip(g0, g1)):
        assert (a - b).abs().max().item() <= 1e-5 * max(a.abs().max().item(), 1e-12), l
    spec_ref = make_grid(16, 2, 14, 16, 2048, mode="reference", seed=2)[0].spec()
    assert ops._auto_perm(spec_ref, x) is None and ops._auto_perm(m.spec(), x) is not None and ops._auto_perm(m.spec(), x[:1000]) is None


@pytest.mark.parametrize("mode,log2T,n", [("reference", 5, 3072), ("reference", 19, 40000), ("trilinear", 5, 3072),
                                          ("trilinear", 14, 20011)])
def test_deterministic_table_gradients(mode, log2T, n):
    """K2d (sorted, single-writer table gradients): bit-identical run to run - also between two different batch
    splits is NOT claimed, only repeatability -, equal to the atomic scatter up to fp32 summation order (rel 1e-5 of max),
    accumulating like it, and equal to the oracle's dense autograd gradient.  T = 2^5 tables: every row collects hundreds of
    contributions (the shipped configs)."""
    from idrk import kernels as K
    from idrk.model.embeddings.hashGridEmbedding import MultiResHashGridMLP
    L, F = (6, 2) if log2T == 5 else (16, 2)
    base, des = (64, 512) if log2T == 5 else (16, 2048)
    gen = torch.Generator().manual_seed(n)
    m = MultiResHashGridMLP(True, 3, L, F, log2T, base, des, frac_mode=mode).to(DEV)
    spec, tables, B = m.spec(), tuple(t.detach() for t in m.tables()), m.freq_encoding.B
    x = (torch.rand(n, 3, generator=gen) * 2 - 1).to(DEV)
    dy = torch.randn(n, K.pad4(spec.width), generator=gen).to(DEV)

    def grads(det, init=0.0):
        gt = [torch.full_like(t, init) for t in tables]
        K.hash_encode_bwd(spec, x, tables, B, dy, gt, False, deterministic=det)
        return gt
    a, b = grads(True), grads(True)
    for l, (p, q) in enumerate(zip(a, b)):
        assert torch.equal(p, q), l                                   # repeatable to the bit
    c = grads(False)
    for l, (p, q) in enumerate(zip(a, c)):
        assert (p - q).abs().max().item() <= 1e-5 * max(q.abs().max().item(), 1e-12), l
    d = grads(True, init=1.5)                                         # accumulates into what is there
    for l, (p, q) in enumerate(zip(a, d)):
        assert (q - 1.5 - p).abs().max().item() <= 2e-6 * max(p.abs().max().item(), 1.0), l
    if n <= 5000:
        sd = {"levels.%d.embedding.weight" % l: t.detach().cpu().clone().requires_grad_(True) for l, t in enumerate(tables)}
        sd["freq_encoding.B"] = B.cpu()
        y = O.hashgrid_embed(x.cpu(), sd, "", L, base, des, mode)
        (y * dy[:, :spec.width].cpu()).sum().backward()
        for l, gtab in enumerate(a):
            ref = sd["levels.%d.embedding.weight" % l].grad
            assert (gtab.cpu() - ref).abs().max().item() <= 1e-5 * max(ref.abs().max().item(), 1e-12), l
    # the module-level switch routes autograd through the same pass
    K.set_deterministic_table_grads(True)
    try:
        outs = []
        for _ in range(2):
            for t in m.tables():
                t.grad = None
            (m(x) * dy[:, :spec.width]).sum().backward()
            outs.append([t.grad.clone() for t in m.tables()])
        for p, q in zip(*outs):
            assert torch.equal(p, q)
        for l, (p, q) in enumerate(zip(outs[0], a)):
            assert torch.equal(p, q), l
    finally:
        K.set_deterministic_table_grads(False)


def test_fourier_prefix_max_error_is_recorded():
    """The Fourier prefix uses a Cody-Waite reduction + the SFU sin / cos (csrc/hash_common.cuh sincos_fast).  Measured
    against float64 on the encoder's real argument range (x in [-1, 1]^3, B ~ N(0, sigma^2) with the reference's sigma
    for base 16 -> 2048, |2 pi x B| up to ~15): the kernel's own error - SFU approximation + the fp32 rounding of the
    argument - must stay below 2e-6 (measured: 1.48e-6 in total, 4.4e-7 from the sin / cos evaluation itself, the rest is
    the fp32 rounding of arguments up to 12.6), which is what leaves room inside the 4e-6 parity bar against the host's fp32
    libm (whose argument is rounded in a different order).  Prints the measured maxima."""
    from idrk import kernels as K
    gen = torch.Generator().manual_seed(11)
    C = 16
    sigma = O.fourier_sigma(16, 2048)
    Bm = torch.randn(3, C, generator=gen) * sigma
    x = torch.rand(1 << 18, 3, generator=gen) * 2 - 1
    spec = K.HashGridSpec([], [], 2, 0, C)
    y = K.hash_encode_fwd(spec, x.to(DEV), (), Bm.to(DEV))[:, :3 + 2 * C].cpu().double()
    xp64 = (2 * np.pi * x.double()) @ Bm.double()
    err_exact = max((y[:, 3:3 + C] - torch.sin(xp64)).abs().max().item(), (y[:, 3 + C:] - torch.cos(xp64)).abs().max().item())
    # the same fp32 argument the kernel forms (x * 2 pi, then a 3-term fma chain), evaluated exactly: isolates the SFU path
    two_pi = np.float32(6.283185307179586)
    xs = (x.numpy() * two_pi).astype(np.float32)
    Bn = Bm.numpy()
    xp32 = (xs[:, 0:1] * Bn[0:1]).astype(np.float32)
    xp32 = (xs[:, 1:2].astype(np.float64) * Bn[1:2] + xp32).astype(np.float32)
    xp32 = (xs[:, 2:3].astype(np.float64) * Bn[2:3] + xp32).astype(np.float32)
    err_sfu = max(np.abs(y[:, 3:3 + C].numpy() - np.sin(xp32.astype(np.float64))).max(),
                  np.abs(y[:, 3 + C:].numpy() - np.cos(xp32.astype(np.float64))).max())
    print("fourier prefix: max |arg| %.2f, max error vs float64 %.3e, of which sin/cos evaluation at the fp32 argument %.3e" % (
        float(xp64.abs().max()), err_exact, err_sfu))
    assert err_sfu <= 6e-7, err_sfu
    assert err_exact <= 2e-6, err_exact


@pytest.mark.parametrize("include_input", [True, False])
def test_posenc_recorded_backward_is_one_differentiable_kernel(include_input):
    """Double backward through the positional encoding (the filter banks inside ImplicitNetwork.gradient with
    create_graph=True): the recorded d/dx is one kernel with its own backward kernel (autograd_ops._PosEncDx); first- and
    second-order results vs the same computation in float64 tensor ops."""
    from idrk import autograd_ops as ops
    gen = torch.Generator().manual_seed(11)
    n, d = 3001, 4
    bands = [2.0 ** k for k in range(6)]
    x0 = (torch.rand(n, d, generator=gen) * 2 - 1)
    width = d * ((2 if include_input else 0) + 2 * len(bands))
    w0 = torch.randn(n, width, generator=gen)

    def run(x, w, enc):
        y = enc(x)
        (gx,) = torch.autograd.grad((y * w).sum(), x, create_graph=True)
        loss = (gx ** 2).sum() + y.sum()
        loss.backward()
        return y.detach(), gx.detach(), x.grad.detach(), w.grad.detach()

    def enc_ref(x):
        cols = [x, x] if include_input else []
        for f in bands:
            cols += [torch.sin(x * f), torch.cos(x * f)]
        return torch.cat(cols, 1)
    xr, wr = x0.double().requires_grad_(True), w0.double().requires_grad_(True)
    ref = run(xr, wr, enc_ref)
    xg, wg = x0.to(DEV).requires_grad_(True), w0.to(DEV).requires_grad_(True)
    got = run(xg, wg, lambda x: ops.positional_encoding(x, bands, include_input))
    for a, b, name in zip(got, ref, ("y", "dy/dx", "d loss / dx", "d loss / dw")):
        scale = b.abs().max().item()
        assert (a.cpu().double() - b).abs().max().item() <= 2e-5 * scale, name
