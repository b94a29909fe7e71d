"""GPU parity of one full IDR training step at the BASELINE shapes (width 512, feature 256), not toy widths.

The product traces its own rays (fp16-pair no-grad pipeline, CUDA tracer); the oracle (fp32 CPU restatement of the
reference, pinned by tests/test_oracle_golden.py) is then fed THAT trace, so both sides differentiate through the same
points and the comparison isolates encode + MLP forward / backward + eikonal double backward + loss at full width.

Two discontinuities are handled by cause, not by loose bounds:

* hit / miss masks: the product's trace vs the oracle's own trace may differ only on rays whose decisive SDF values came
  within the SDF tolerance of a decision boundary (RayTracerOracle.margin);
* ReLU patterns of the rendering network: its inputs (normals, features) agree to ~1e-5, so a handful of pre-activations
  with |z| <= 1e-4 land on the other side of 0 and toggle their whole gradient contribution.  The product's activation
  pattern is tapped (mlp.ACT_PATTERN_TAP), asserted to differ from the oracle's own pattern ONLY at such borderline
  pre-activations, and the oracle's gradients are then taken under the product's pattern - a smooth comparison.

Bars: sdf_output abs 5e-5, rgb abs 2e-4 (filter banks 2e-3), grad_theta 2e-4 of max, losses rel 5e-4, EVERY parameter
gradient within 1e-3 of its max-abs.  For the filter-bank encoders (sin(30 x) / sin(240 x) chains) the fp32 reference
arithmetic is itself only conditionally accurate: the oracle is run a second time in float64, and a quantity's bar is
max(the fixed bar, 3 x the distance between the fp32 and the fp64 oracle) - measured, not guessed; the product is
compared with the float64 result.

Cases: cfg2 exactly as benched (HashGrid L=6 T=2^5, 2048 rays); HashGrid L=16 T=2^19 (cfg1 tables inside the step);
cfg3 (FFB L=6, FFB view embedder); cfg4 shape (StyleModNFFB with 2^22-row tables); FFB L=16 (W = 136 filter bank).
"""
import pytest
import torch

from conftest import assert_flips_borderline
from oracle import idr_oracle as O
from tests_support import RAY_TRACER_CONF, make_conf, quiet_build

pytestmark = pytest.mark.gpu
DEV = "cuda"

#        tag      embed (type, L, log2T, base, desired, bound)        view embed                rays  flip tol
CASES = {
    "cfg2": (("HashGrid", 6, 5, 64, 512, 1.0), None, 2048, 2e-5),
    "hash16_T19": (("HashGrid", 16, 19, 16, 2048, 1.0), None, 1024, 2e-5),
    "cfg3_ffb": (("FFB", 6, 5, 16, 512, 0.45), ("FFB", 4, 3, 16, 512, 1.0), 1024, 1e-4),
    "cfg4_style_T22": (("StyleModNFFB", 6, 22, 16, 512, 0.45), ("StyleModNFFB", 4, 3, 16, 512, 1.0), 1024, 1e-4),
    "ffb16": (("FFB", 16, 14, 16, 2048, 0.45), None, 512, 1e-4),
}
Z_BORDER = 1e-4          # |pre-activation| below which a ReLU unit may legitimately sit on either side


def _build(tag):
    from idrk.model.implicit_differentiable_renderer import IDRNetwork
    emb, view, rays, ftol = CASES[tag]
    et, L, log2T, base, des, bound = emb
    conf = make_conf(et, L, log2T, base, des, bound, view_type=view[0] if view else "NerfPos")
    torch.manual_seed(0)
    model = quiet_build(IDRNetwork, conf)
    # reference init + a perturbation: the geometric init makes all 257 outputs of the last layer nearly identical
    #  and the tables 1e-4-small; perturbed weights make every gradient path carry signal
    gen = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if "embedding.weight" in name:
                p.mul_(300.0)
            elif p.dim() == 2 and p.shape[1] > 1 and "ff_lin" not in name and "out_layer" not in name:
                p.add_(torch.randn(p.shape, generator=gen) * 0.02 * p.abs().mean())
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    cfg = O.IDRCfg(O.EmbedCfg(et, L, log2T, 2, base, des, bound),
                   view_embed=O.EmbedCfg(view[0], view[1], view[2], 2, view[3], view[4], view[5]) if view else None,
                   ray_tracer=dict(RAY_TRACER_CONF))
    return model.to(DEV).train(), sd, cfg, rays, ftol


def _oracle_step(sd, cfg, inp, rgb, eik, u, tr_out, dtype, relu_masks=None):
    """Oracle forward + loss + parameter gradients in `dtype`; `relu_masks` forces the rendering network's activation
    pattern (one [n_surface, width] 0/1 tensor per hidden layer).  Returns (out, losses, grads, relu pre-activations)."""
    zs = []
    masks = list(relu_masks) if relu_masks is not None else None
    real_relu = torch.relu

    def tapped_relu(h):
        zs.append(h.detach().clone())
        if masks is None:
            return real_relu(h)
        return h * masks.pop(0).to(h.dtype)
    torch.set_default_dtype(dtype)
    torch.relu = tapped_relu
    try:
        s = {k: (v.detach().to(dtype) if v.is_floating_point() else v.detach().clone()) for k, v in sd.items()}
        for k, v in s.items():
            if v.is_floating_point() and not k.endswith(".B") and not k.endswith("dencity_net.beta"):
                v.requires_grad_(True)
        cast = lambda t: t.to(dtype) if t.is_floating_point() else t       # noqa: E731
        out = O.idr_forward({k: cast(v) for k, v in inp.items()}, s, cfg, True, cast(eik), cast(u),
                            tracer_out=tuple(cast(t) for t in tr_out))
        lo = O.idr_loss(out, cast(rgb))
        names = [k for k, v in s.items() if v.requires_grad]
        grads = dict(zip(names, torch.autograd.grad(lo["loss"], [s[k] for k in names], allow_unused=True)))
    finally:
        torch.relu = real_relu
        torch.set_default_dtype(torch.float32)
    return out, lo, grads, zs


@pytest.mark.parametrize("tag", list(CASES))
def test_idr_step_full_width_given_same_trace(tag):
    from idrk import mlp
    from idrk.model.loss import IDRLoss
    model, sd, cfg, rays, ftol = _build(tag)
    nffb = cfg.embed.embed_type != "HashGrid"
    inp, rgb = O.synthetic_batch(rays, seed=1)
    gen = torch.Generator().manual_seed(2)
    eik = torch.rand(rays // 2, 3, generator=gen) * 2 - 1
    u = torch.rand(100, generator=gen)
    model.injected_eikonal_points, model.ray_tracer.injected_min_sdf_steps = eik, u
    traced = model.trace({k: v.to(DEV) for k, v in inp.items()})
    tap = []
    mlp.ACT_PATTERN_TAP[0] = tap
    try:
        out = model.shade(traced)
    finally:
        mlp.ACT_PATTERN_TAP[0] = None
    lo = IDRLoss(0.1, 100.0, 50.0)(out, {"rgb": rgb.to(DEV)})
    lo["loss"].backward()

    tr_out = (out["points"].detach().cpu(), traced["network_object_mask"].cpu(), traced["dists"].cpu())
    surf = (traced["network_object_mask"] & traced["object_mask"]).cpu()
    n_s = int(surf.sum())
    assert n_s >= 20, "only %d surface rays: the rendering path would hardly be exercised" % n_s

    # (1) the oracle's own run: values, and its rendering-network pre-activations
    o_own, lo_own, _, zs = _oracle_step(sd, cfg, inp, rgb, eik, u, tr_out, torch.float32)
    assert len(tap) == len(zs) == cfg.n_lin_rgb - 1
    z_border = Z_BORDER
    if nffb:
        # how far the fp32 reference arithmetic itself is from exact on these pre-activations (sin(w0 .) chains)
        zs64 = _oracle_step(sd, cfg, inp, rgb, eik, u, tr_out, torch.float64)[3]
        z_border = max(Z_BORDER, 3.0 * max((a.double() - b).abs().max().item() for a, b in zip(zs, zs64)))
    relu_masks, n_toggled = [], 0
    for S, z in zip(tap, zs):
        m = (S[:, :z.shape[1]].cpu()[surf] > 0)
        differs = m != (z > 0)
        n_toggled += int(differs.sum())
        assert (z[differs].abs() <= z_border).all(), "a ReLU unit with |z| = %.3g sits on the other side (border %.3g)" % (
            float(z[differs].abs().max()), z_border)
        relu_masks.append(m.float())
    assert n_toggled <= max(8, int(2e-3 * n_s * 512 * len(zs))), n_toggled

    # (2) the oracle under the product's activation pattern, fp32 and (filter banks) float64
    o32, lo32, g32, _ = _oracle_step(sd, cfg, inp, rgb, eik, u, tr_out, torch.float32, relu_masks)
    if nffb:
        o_ref, lo_ref, g_ref, _ = _oracle_step(sd, cfg, inp, rgb, eik, u, tr_out, torch.float64, relu_masks)
    else:
        o_ref, lo_ref, g_ref = o32, lo32, g32

    def rel_max(a, b):
        return (a.double() - b.double()).abs().max().item() / max(b.abs().max().item(), 1e-30)

    report = ["surface rays %d, ReLU units toggled at |z| <= %.1e: %d" % (n_s, z_border, n_toggled)]

    def check(name, err, tol, cond=0.0):
        tol_eff = max(tol, 3.0 * cond)
        report.append("%-84s %.3e (bar %.1e%s)%s" % (name, err, tol_eff, ", fp32 vs fp64 oracle %.1e" % cond if cond else "",
                                                    "  <-- FAIL" if not err <= tol_eff else ""))
        return err <= tol_eff

    ok = True
    c = lambda k: (o32[k].double() - o_ref[k].double()).abs().max().item() if nffb else 0.0      # noqa: E731
    ok &= check("sdf_output abs", (out["sdf_output"].cpu().double() - o_ref["sdf_output"].double()).abs().max().item(), 5e-5, c("sdf_output"))
    ok &= check("rgb_values abs", (out["rgb_values"].cpu().double() - o_ref["rgb_values"].double()).abs().max().item(),
                2e-3 if nffb else 2e-4, c("rgb_values"))
    ok &= check("grad_theta / max", rel_max(out["grad_theta"].cpu(), o_ref["grad_theta"]), 2e-4,
                rel_max(o32["grad_theta"], o_ref["grad_theta"]) if nffb else 0.0)
    for k in ("loss", "rgb_loss", "eikonal_loss", "mask_loss"):
        ref = float(lo_ref[k])
        ok &= check(k + " rel", abs(float(lo[k]) - ref) / max(1.0, abs(ref)), 5e-4,
                    abs(float(lo32[k]) - ref) / max(1.0, abs(ref)) if nffb else 0.0)
        # the pattern swap changes the loss values only through units with |z| <= z_border
        assert abs(float(lo_own[k]) - float(lo32[k])) <= max(1e-5, z_border) * max(1.0, abs(float(lo32[k]))), k
    pd = dict(model.named_parameters())
    n_checked = 0
    for k, gq in g_ref.items():
        p = pd[k]
        if gq is None or gq.abs().max() == 0:
            ok &= check(k + " (zero grad) abs", 0.0 if p.grad is None else p.grad.abs().max().item(), 1e-7)
            continue
        n_checked += 1
        ok &= check(k + " grad / max", rel_max(p.grad.cpu(), gq), 1e-3, rel_max(g32[k], gq) if nffb else 0.0)
    assert n_checked >= 30, n_checked
    assert ok, "\n" + "\n".join(report)
    print("\n".join(report))

    # (3) the trace itself against the oracle's own trace: flips only on borderline rays
    orc = O.RayTracerOracle(**RAY_TRACER_CONF)
    dirs_o, cam_o = O.camera_rays(inp["uv"], inp["pose"], inp["intrinsics"])
    sd_ng = {k: v.detach() for k, v in sd.items()}
    with torch.no_grad():
        p_o, m_o, d_o = orc(lambda q: O.implicit_forward(q, sd_ng, cfg.embed)[:, 0], cam_o, inp["object_mask"].reshape(-1),
                            dirs_o, u)
    m = traced["network_object_mask"].cpu()
    flips = assert_flips_borderline(m, m_o, orc.margin, ftol, tag)
    assert flips <= max(2, rays // 100), flips
