"""GPU parity of ImplicitNetwork / RenderingNetwork (forward, gradient(), eikonal double backward) against
the golden vectors of the real reference and against the oracle at full width.
Tolerance (3xTF32 tensor-core mode, fp32-accurate): rel 1e-4 of max-abs + abs 2e-5 on outputs,
2e-3 of max-abs on parameter gradients of the second-order loss."""
import numpy as np
import pytest
import torch

from conftest import sd_from
from oracle import idr_oracle as O
from tests_support import load_sd_into, make_conf, quiet_build

pytestmark = pytest.mark.gpu
DEV = "cuda"

NET_CFGS = {"hash": ("HashGrid", 6, 5, 64, 512, 1.0), "hash16": ("HashGrid", 16, 8, 16, 2048, 1.0),
            "ffb": ("FFB", 6, 5, 16, 512, 0.45), "style": ("StyleModNFFB", 6, 5, 16, 512, 0.45)}


def T(a):
    return torch.from_numpy(np.array(a))


def close(a, b, rel=1e-4, abs_=2e-5):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return (a - b).abs().max().item() <= rel * b.abs().max().item() + abs_


def build_model(tag, g):
    from idrk.model.implicit_differentiable_renderer import IDRNetwork
    et, L, log2T, base, des, bound = NET_CFGS[tag]
    conf = make_conf(et, L, log2T, base, des, bound, width=96, feature=32)
    model = quiet_build(IDRNetwork, conf)
    load_sd_into(model, sd_from(g, "sd_%s/" % tag))
    return model.to(DEV)


@pytest.mark.parametrize("tag", ["hash", "hash16", "ffb", "style"])
def test_implicit_network_golden(golden, tag):
    g = golden("networks")
    model = build_model(tag, g)
    net = model.implicit_network
    x = T(g["x"]).to(DEV)
    y = net(x)
    assert close(y, T(g["y_" + tag]))
    with torch.no_grad():
        y_inf = net(x)
        s_inf = net.sdf(x)
    assert close(y_inf, T(g["y_" + tag]))
    assert close(s_inf, T(g["y_" + tag])[:, 0])
    gr = net.gradient(x.clone())
    nffb = tag in ("ffb", "style")       # sin(30 x) chains amplify last-ulp differences of sin/cos
    assert close(gr, T(g["grad_" + tag]), rel=1e-3 if nffb else 2e-4)
    loss = ((gr[:, 0, :].norm(2, dim=1) - 1) ** 2).mean() + y[:, 0].mean() + 0.01 * (y[:, 1:] ** 2).mean()
    assert abs(loss.item() - float(g["loss_" + tag][0])) <= 2e-4 * max(1.0, abs(float(g["loss_" + tag][0])))
    names = [k for k in g if k.startswith("pg_%s/" % tag)]
    pdict = dict(net.named_parameters())
    params = [pdict[k.split("/", 1)[1]] for k in names]
    grads = torch.autograd.grad(loss, params, allow_unused=True)
    for k, gq in zip(names, grads):
        ref = T(g[k])
        if gq is None:
            assert ref.abs().max() == 0, k
            continue
        assert close(gq, ref, rel=1e-2 if nffb else 2e-3, abs_=1e-7), k


@pytest.mark.parametrize("tag", ["hash", "hash16", "ffb", "style", "viewffb"])
def test_rendering_network_golden(golden, tag):
    from idrk.model.implicit_differentiable_renderer import IDRNetwork
    g = golden("networks")
    if tag == "viewffb":
        conf = make_conf("FFB", 6, 5, 16, 512, 0.45, view_type="FFB", width=96, feature=32)
        model = quiet_build(IDRNetwork, conf)
        load_sd_into(model, sd_from(g, "sd_viewffb/"))
        model = model.to(DEV)
    else:
        model = build_model(tag, g)
    inp = T(g["rn_in_" + tag]).to(DEV)
    rgb = model.rendering_network(inp[:, 0:3], inp[:, 3:6], inp[:, 6:9], inp[:, 9:])
    assert close(rgb, T(g["rn_out_" + tag]))


@pytest.mark.parametrize("prec,rel", [("3xtf32", 1e-4), ("tf32", 2e-2), ("fp32", 1e-4)])
def test_full_width_vs_oracle(prec, rel):
    """BASELINE cfg1: MultiResHash (16 levels, 2^19, F=2) + ImplicitNetwork 8x512, 8192 points."""
    from idrk import kernels as K
    from idrk.model.implicit_differentiable_renderer import ImplicitNetwork
    cfg = O.EmbedCfg("HashGrid", 16, 19, 2, 16, 2048, 1.0)
    gen = torch.Generator().manual_seed(0)
    sd = O.make_implicit_sd(cfg, gen, perturb=0.01)
    for k in list(sd):
        if "embedding.weight" in k:
            sd[k] = sd[k] * 1000.0
    net = quiet_build(ImplicitNetwork, 256, 3, 1, [512] * 8, True, 0.6, (4,), True, 16, "HashGrid", 19, 2, 16, 2048, 1.0)
    load_sd_into(net, sd, "implicit_network.")
    net = net.to(DEV)
    x = torch.rand(8192, 3, generator=torch.Generator().manual_seed(0)) * 2 - 1
    K.set_precision(prec)
    try:
        y = net(x.to(DEV))
        y[:, 0].sum().backward()
        with torch.no_grad():
            s = net.sdf(x.to(DEV))
    finally:
        K.set_precision("3xtf32")
    for v in sd.values():
        if v.dtype == torch.float32 and v.dim() > 0 and not (v.dim() == 2 and v.shape[0] == 3):
            v.requires_grad_(True)
    ref = O.implicit_forward(x, sd, cfg)
    ref[:, 0].sum().backward()
    assert close(y, ref, rel=rel)
    assert close(s, ref[:, 0], rel=rel)
    assert close(net.lin3.weight_v.grad, sd["implicit_network.lin3.weight_v"].grad, rel=max(rel, 1e-3), abs_=1e-7)
    assert close(net.lin0.weight_g.grad, sd["implicit_network.lin0.weight_g"].grad, rel=max(rel, 1e-3), abs_=1e-7)
    g_ref = sd["implicit_network.embed_model.embedder_obj.levels.5.embedding.weight"].grad
    g_got = net.embed_model.embedder_obj.levels[5].embedding.weight.grad
    assert close(g_got, g_ref, rel=max(rel, 1e-3), abs_=1e-8)


@pytest.mark.parametrize("tag", ["ffb", "style"])
def test_fused_filter_bank_encoder_matches_module_and_golden(golden, tag):
    """csrc/nffb.cu (one launch, FP32 FMAs) vs the module path (contraction / posenc kernels) vs the reference's
    golden embedding-dependent outputs; plus the device-side row count (rows beyond it are not written)."""
    from idrk import kernels as K
    g = golden("networks")
    model = build_model(tag, g)
    net = model.implicit_network
    ffb = net.embed_model.embedder_obj
    assert net._fused_filter_bank() is ffb
    x = T(g["x"]).to(DEV)
    with torch.no_grad():
        ref = ffb(x)
        got = K.nffb_encode_fwd(ffb, x)
    assert got.shape[1] == K.pad4(ffb.embeddings_dim)
    # sin(w0 .) chains with the reference's SIREN weights amplify last-ulp differences of the two summation orders
    assert close(got[:, :ffb.embeddings_dim], ref, rel=2e-4, abs_=2e-6)
    assert (got[:, ffb.embeddings_dim:] == 0).all()
    xs = torch.rand(3001, 3, generator=torch.Generator().manual_seed(3)).to(DEV) * 0.8 - 0.4
    cnt = torch.tensor([1234], device=DEV, dtype=torch.int32)
    out = torch.full((3001, K.pad4(ffb.embeddings_dim)), 7.0, device=DEV)
    with torch.no_grad():
        K.nffb_encode_fwd(ffb, xs, out=out, m_count=cnt)
        full = ffb(xs)
    assert close(out[:1234, :ffb.embeddings_dim], full[:1234], rel=2e-4, abs_=2e-6)
    assert (out[1234:] == 7.0).all()
    # pair-output form (the SDF pipeline's operand, tensor-core kernel): h + l / 2^11 == the fp32 row to fp16-pair rounding,
    # the scaled second copy lands in its column slot, pads are zero, rows beyond the count untouched
    if K.nffb_pair_supported(ffb):
        E = ffb.embeddings_dim
        ld, ld2, off = K.pad8(E), K.pad8(100 + E), 100
        h, l = (torch.full((3001, ld), 3.0, device=DEV, dtype=torch.float16) for _ in range(2))
        h2, l2 = (torch.full((3001, ld2), 3.0, device=DEV, dtype=torch.float16) for _ in range(2))
        K.nffb_encode_f16pair(ffb, xs, 3001, h, l, ld, ld - E, cnt, second=(h2[:, off:], l2[:, off:], ld2, ld2 - off - E, 0.5))
        ref32 = out[:1234, :E].double()
        got1 = h[:1234, :E].double() + l[:1234, :E].double() / 2048.0
        got2 = h2[:1234, off:off + E].double() + l2[:1234, off:off + E].double() / 2048.0
        assert (got1 - ref32).abs().max().item() <= 4e-7 * max(1.0, ref32.abs().max().item())
        assert (got2 - 0.5 * ref32).abs().max().item() <= 4e-7 * max(1.0, ref32.abs().max().item())
        assert (h[:1234, E:] == 0).all() and (l[:1234, E:] == 0).all() and (h2[:1234, off + E:] == 0).all()
        assert (h[1234:] == 3.0).all() and (h2[1234:] == 3.0).all() and (h2[:, :off] == 3.0).all()
