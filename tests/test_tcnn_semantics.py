"""HashGridTcnn / FFBTcnn selector entries without tiny-cuda-nn.

Parity is UNPINNED here: tiny-cuda-nn is an unpinned, un-vendored dependency of the reference
(tcnn_src/hashGridEncoderTcnn.py:63-80) and cannot be run in this environment.  The oracle
(oracle/idr_oracle.py::ngp_grid_encode) restates its published Grid/Hash/Linear algorithm; the CPU tests below pin
that restatement with interpolation identities that do not depend on it being bug-compatible with itself:
a dense level whose vertex values are an affine function of the vertex position must reproduce that function of
pos = x * scale + 0.5 exactly (checks vertex order, strides, the +0.5 offset and the weights), and hashed rows
must come from the three published primes.  GPU tests compare the kernel mode IDRK_HASH_NGP with the oracle:
corner rows bit-exact, features abs 2e-6, table gradients rel 1e-5 of max, dL/dx rel 1e-3 away from cell faces.
"""
import numpy as np
import pytest
import torch

from oracle import idr_oracle as O

DEV = "cuda"


# ----------------------------------------------------------------------------------------------
# CPU: oracle identities, layout, selector plumbing
# ----------------------------------------------------------------------------------------------
def test_ngp_layout_matches_published_rules():
    scales, res, rows, offs = O.ngp_level_layout(6, 12, 16, 2.0)
    assert scales == [15.0, 31.0, 63.0, 127.0, 255.0, 511.0]
    assert res == [16, 32, 64, 128, 256, 512]
    assert rows == [4096] * 6 and offs == [0, 4096, 8192, 12288, 16384, 20480, 24576]
    scales, res, rows, _ = O.ngp_level_layout(4, 19, 16, 1.5)
    assert res == [16, 24, 36, 54] and np.allclose(scales, [15.0, 23.0, 35.0, 53.0])
    assert rows == [4096, 13824, 46656, 157464]                    # all dense: R^3 (multiples of 8) below 2^19
    _, res, rows, _ = O.ngp_level_layout(3, 19, 15, 1.3)
    assert rows[0] == (res[0] ** 3 + 7) // 8 * 8                   # rounded up to a multiple of 8


def test_ngp_dense_level_reproduces_affine_field():
    """Vertex (i, j, k) of a dense level stores a + b . (i, j, k): trilinear interpolation must return
    a + b . pos with pos = x * scale + 0.5, wherever the +1 corner is still a vertex of the level (for x close to 1 the
    +1 vertex has index R and aliases into the next row of the table: tiny-cuda-nn's behaviour, kept)."""
    L, F, log2T, base = 2, 2, 19, 8
    scales, res, rows, offs = O.ngp_level_layout(L, log2T, base, 2.0)
    assert all(r ** 3 <= n for r, n in zip(res, rows))
    gen = torch.Generator().manual_seed(3)
    params = torch.zeros(offs[-1] * F)
    coef = []
    for l in range(L):
        R = res[l]
        a, b = torch.randn(F, generator=gen), torch.randn(F, 3, generator=gen)
        i = torch.arange(R ** 3)
        v = torch.stack([i % R, (i // R) % R, i // (R * R)], -1).float()              # x fastest, then y, then z
        params[offs[l] * F: offs[l] * F + R ** 3 * F] = (a + v @ b.t()).reshape(-1)
        coef.append((a, b))
    x = torch.rand(4000, 3, generator=gen) * 0.78        # floor(pos) + 1 <= R - 1: no wrap at the upper faces
    y = O.ngp_grid_encode(x, params, L, F, log2T, base, 2.0)
    for l in range(L):
        a, b = coef[l]
        pos = x * scales[l] + 0.5
        assert torch.allclose(y[:, l * F:(l + 1) * F], a + pos @ b.t(), atol=2e-4)


def test_ngp_hashed_level_uses_published_primes():
    L, log2T, base = 3, 10, 16
    scales, res, rows, _ = O.ngp_level_layout(L, log2T, base, 2.0)
    assert rows[2] == 1024 and res[2] ** 3 > rows[2]
    x = torch.tensor([[0.3, 0.6, 0.9], [0.01, 0.5, 0.25]])
    idx, fr = O.ngp_corner_rows(x, scales[2], res[2], rows[2])
    for p in range(2):
        pos = x[p].double().numpy() * scales[2] + 0.5
        g = np.floor(pos).astype(np.int64)
        assert np.allclose(fr[p].numpy(), pos - g, atol=1e-5)
        for k in range(8):
            c = [int(g[d]) + ((k >> d) & 1) for d in range(3)]
            h = (c[0] ^ ((c[1] * 2654435761) & 0xFFFFFFFF) ^ ((c[2] * 805459861) & 0xFFFFFFFF)) % 1024
            assert idx[p, k] == h


def test_selector_builds_tcnn_entries():
    from idrk.model.custom_embedder_decoder import Custom_Embedding_Network
    m = Custom_Embedding_Network(3, [3, 512], 'HashGridTcnn', 6, 12, 2, 16, 512, 1.0)
    assert m.embeddings_dim == 3 + 6 * 2
    assert list(m.state_dict().keys()) == ['embedder_obj.grid_encoder.params']          # tcnn's key
    g = m.embedder_obj.grid_encoder
    assert (g.scales, g.resolutions, g.rows, g.offsets) == O.ngp_level_layout(6, 12, 16, 2.0)
    assert g.params.numel() == 6 * 4096 * 2 and float(g.params.abs().max()) <= 1e-4
    f = Custom_Embedding_Network(3, [3, 512], 'FFBTcnn', 6, 12, 2, 16, 512, 0.45)
    W = 2 * (2 + 2 * 6)                                # PositionalEncoding(input_dims=2, L freqs) width, not doubled
    assert f.embeddings_dim == 3 + W and f.embedder_obj.ff_lin1.weight.shape == (W, W)
    assert hasattr(f.embedder_obj, "StyleAttentionBlock")
    with pytest.raises(ValueError):
        Custom_Embedding_Network(3, [3, 512], 'HashGridCUDA', 6, 12, 2, 16, 512, 1.0)


# ----------------------------------------------------------------------------------------------
# GPU: kernel mode IDRK_HASH_NGP vs the oracle
# ----------------------------------------------------------------------------------------------
def _grid(L, log2T, base, pls, seed):
    from idrk.model.embeddings.tcnn_src.hashGridEncoderTcnn import NgpGrid
    g = NgpGrid(L, 2, log2T, base, pls)
    gen = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        g.params.copy_(torch.randn(g.params.shape, generator=gen) * 0.5)
    return g.to(DEV)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [(6, 12, 16, 2.0, 4096 + 5), (16, 15, 16, 1.38, 3001), (5, 19, 8, 1.7, 1000), (8, 14, 4, 2.0, 64)])
def test_ngp_forward_backward_vs_oracle(cfg):
    from idrk import kernels as K
    L, log2T, base, pls, n = cfg
    g = _grid(L, log2T, base, pls, seed=L)
    gen = torch.Generator().manual_seed(n)
    x = torch.rand(n, 3, generator=gen)
    w = torch.randn(n, 2 * L, generator=gen)
    # corner rows: bit exact
    _, idx = K.hash_encode_fwd(g.spec(), x.to(DEV), g.tables(), None, want_idx=True)
    scales, res, rows, offs = O.ngp_level_layout(L, log2T, base, pls)
    assert (g.scales, g.resolutions, g.rows, g.offsets) == (scales, res, rows, offs)
    assert any(r ** 3 <= m for r, m in zip(res, rows)) and (L < 6 or any(r ** 3 > m for r, m in zip(res, rows)))
    for l in range(L):
        ref_idx, _ = O.ngp_corner_rows(x, scales[l], res[l], rows[l])
        assert np.array_equal(idx[:, l, :].cpu().numpy().astype(np.int64) & 0xFFFFFFFF, ref_idx), l
    # forward + gradients through autograd
    xd = x.to(DEV).requires_grad_(True)
    y = g(xd)
    (y * w.to(DEV)).sum().backward()
    params = g.params.detach().cpu().clone().requires_grad_(True)
    ref = O.ngp_grid_encode(x, params, L, 2, log2T, base, pls)
    (ref * w).sum().backward()
    assert torch.allclose(y.detach().cpu(), ref.detach(), atol=2e-6, rtol=1e-5)
    gp, gr = g.params.grad.cpu(), params.grad
    assert torch.allclose(gp, gr, atol=1e-5 * gr.abs().max().item(), rtol=1e-4)
    dx_ref = O.ngp_grid_dx(x, params, w, L, 2, log2T, base, pls)
    bad = ~torch.isclose(xd.grad.cpu(), dx_ref, atol=1e-3 * dx_ref.abs().max().item(), rtol=1e-3)
    assert bad.sum().item() <= 3                       # dL/dx is discontinuous at cell faces


@pytest.mark.gpu
def test_ngp_dense_level_affine_field_on_gpu():
    """The interpolation identity of the CPU test, through the kernel."""
    from idrk.model.embeddings.tcnn_src.hashGridEncoderTcnn import NgpGrid
    g = NgpGrid(2, 2, 19, 8, 2.0)
    gen = torch.Generator().manual_seed(5)
    coef = []
    with torch.no_grad():
        g.params.zero_()
        for l in range(2):
            R = g.resolutions[l]
            a, b = torch.randn(2, generator=gen), torch.randn(2, 3, generator=gen)
            i = torch.arange(R ** 3)
            v = torch.stack([i % R, (i // R) % R, i // (R * R)], -1).float()
            g.params[g.offsets[l] * 2: g.offsets[l] * 2 + R ** 3 * 2] = (a + v @ b.t()).reshape(-1)
            coef.append((a, b))
    g = g.to(DEV)
    x = torch.rand(5000, 3, generator=gen) * 0.78
    y = g(x.to(DEV)).cpu()
    for l in range(2):
        a, b = coef[l]
        assert torch.allclose(y[:, 2 * l:2 * l + 2], a + (x * g.scales[l] + 0.5) @ b.t(), atol=2e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("embed_type", ["HashGridTcnn", "FFBTcnn"])
def test_implicit_network_with_tcnn_entries(embed_type):
    """The selector entries inside ImplicitNetwork: forward, gradient() and a loss on the gradient (double backward) run,
    the no-grad SDF query equals the autograd forward's column 0, and the traced mask path works (sync-free tracer)."""
    from idrk.model.implicit_differentiable_renderer import ImplicitNetwork
    from tests_support import quiet_build
    torch.manual_seed(0)
    net = quiet_build(ImplicitNetwork, 32, d_in=3, d_out=1, dims=[96] * 4, geometric_init=True, bias=0.6, skip_in=[2],
                      weight_norm=True, multires=6, embed_type=embed_type, log2_max_hash_size=12, max_points_per_entry=2,
                      base_resolution=16, desired_resolution=512, bound=1.0).to(DEV)
    with torch.no_grad():           # the geometric init zeroes the embedding columns of lin0: give them weight
        net.lin0.weight_v.add_(torch.randn(net.lin0.weight_v.shape, generator=torch.Generator().manual_seed(2)).to(DEV) * 0.2)
        p0 = net.embed_model.embedder_obj
        (p0.grid_encoder if embed_type == "HashGridTcnn" else p0.grid_enc.grid_encoder).params.mul_(1000.0)
    x = (torch.rand(777, 3, generator=torch.Generator().manual_seed(1)) * 0.9 + 0.05).to(DEV)
    out = net(x)
    assert out.shape == (777, 33) and torch.isfinite(out).all()
    g = net.gradient(x.clone())
    loss = ((g.norm(2, dim=-1) - 1) ** 2).mean() + out[:, 0].mean()
    loss.backward()
    p = net.embed_model.embedder_obj
    params = p.grid_encoder.params if embed_type == "HashGridTcnn" else p.grid_enc.grid_encoder.params
    assert params.grad is not None and torch.isfinite(params.grad).all() and params.grad.abs().max() > 0
    with torch.no_grad():
        s = net.sdf(x)
    assert (s - out[:, 0].detach()).abs().max().item() <= 5e-5
