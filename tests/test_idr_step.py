"""GPU parity of a full IDR training step (ray trace + encode + MLP fwd/bwd + eikonal + loss backward)
against the golden vectors of the real reference (small model, 256 rays) and the oracle.
Tolerances: masks equal up to <= 1 % borderline rays; on rays whose masks agree points/sdf abs 2e-4,
rgb abs 2e-3, losses rel 2e-3, dense parameter gradients within 8 % of max-abs (a handful of rays land in a
neighbouring hash cell than in the fixture; test_differentiable_part_given_same_trace pins the same gradients to 1 %)."""
import numpy as np
import pytest
import torch

from conftest import assert_flips_borderline, sd_from
from tests_support import load_sd_into, make_conf, quiet_build

pytestmark = pytest.mark.gpu
DEV = "cuda"
IDR_CFGS = {"hash": ("HashGrid", 6, 5, 64, 512, 1.0), "style": ("StyleModNFFB", 6, 5, 16, 512, 0.45)}


def T(a):
    return torch.from_numpy(np.array(a))


@pytest.mark.parametrize("tag", list(IDR_CFGS))
def test_idr_step_golden(golden, tag):
    from idrk.model.implicit_differentiable_renderer import IDRNetwork
    from idrk.model.loss import IDRLoss
    g = golden("idr_step")
    et, L, log2T, base, des, bound = IDR_CFGS[tag]
    model = quiet_build(IDRNetwork, make_conf(et, L, log2T, base, des, bound, width=96, feature=32))
    load_sd_into(model, sd_from(g, "sd_%s/" % tag))
    model = model.to(DEV).train()
    model.injected_eikonal_points = T(g["eik_points_" + tag])
    model.ray_tracer.injected_min_sdf_steps = T(g["min_sdf_steps_" + tag])
    inp = {"uv": T(g["uv"]).to(DEV), "pose": T(g["pose"]).to(DEV), "intrinsics": T(g["K"]).to(DEV),
           "object_mask": T(g["mask"]).to(DEV)}
    out = model(inp)
    lo = IDRLoss(eikonal_weight=0.1, mask_weight=100.0, alpha=50.0)(out, {"rgb": T(g["rgb_gt"]).to(DEV)})
    lo["loss"].backward()

    m_ref = T(g["network_object_mask_" + tag])
    m = out["network_object_mask"].cpu()
    flips = (m != m_ref).sum().item()
    assert flips <= max(1, m.numel() // 100), flips
    if flips:
        # count AND cause: the oracle (== the reference fixture on these inputs, tests/test_oracle_golden.py) records how
        # close every ray came to deciding differently; the fp16-pair SDF agrees with fp32 to 1e-5, so only rays within
        # 2e-5 (NFFB: sin(30 x) chains, 1e-4) of a decision boundary may flip
        from oracle import idr_oracle as O
        from tests_support import RAY_TRACER_CONF
        cfg = O.EmbedCfg(et, L, log2T, 2, base, des, bound)
        sd_o = sd_from(g, "sd_%s/" % tag)
        orc = O.RayTracerOracle(**RAY_TRACER_CONF)
        dirs_o, cam_o = O.camera_rays(T(g["uv"]), T(g["pose"]), T(g["K"]))
        with torch.no_grad():
            _, m_orc, _ = orc(lambda q: O.implicit_forward(q, sd_o, cfg)[:, 0], cam_o, T(g["mask"]).reshape(-1), dirs_o,
                              T(g["min_sdf_steps_" + tag]))
        assert torch.equal(m_orc, m_ref.reshape(-1))
        assert_flips_borderline(m, m_ref, orc.margin, 1e-4 if tag == "style" else 2e-5, tag)
    agree = m == m_ref
    dp = (out["points"].cpu() - T(g["points_" + tag])).abs().max(1).values
    assert (dp[agree] > 2e-4).sum().item() <= max(2, m.numel() // 20)     # argmin ties of the 100-sample sweeps
    ok = agree & (dp <= 5e-6)          # rays traced to (numerically) the same point
    assert ok.sum().item() >= m.numel() // 2
    # the hash feature is piecewise constant: a 1e-6 shift of a point across a cell boundary moves its SDF by a
    # table-value difference, so a few outliers are legitimate; bound their count instead of the maximum
    lim = max(2, m.numel() // 16)
    d_sdf = (out["sdf_output"].cpu() - T(g["sdf_output_" + tag])).abs().squeeze(-1)[ok]
    assert (d_sdf > 2e-4).sum().item() <= lim
    d_rgb = (out["rgb_values"].cpu() - T(g["rgb_values_" + tag])).abs().max(1).values[ok]
    assert (d_rgb > 2e-3).sum().item() <= lim
    nffb = tag == "style"
    if flips == 0:
        gt_ref = T(g["grad_theta_" + tag])
        tol = (2e-3 if nffb else 3e-4) * gt_ref.abs().max().item()
        n_eik = g["eik_points_" + tag].shape[0]
        assert (out["grad_theta"].cpu()[:n_eik] - gt_ref[:n_eik]).abs().max().item() <= tol
    for k in ("loss", "rgb_loss", "eikonal_loss", "mask_loss"):
        ref = float(g["%s_%s" % (k, tag)][0])
        assert abs(float(lo[k]) - ref) <= (2e-2 if flips else 2e-3) * max(1.0, abs(ref)), (k, float(lo[k]), ref)
    pd = dict(model.named_parameters())
    worst = 0.0
    for k in g:
        if not k.startswith("pg_%s/" % tag):
            continue
        ref = T(g[k])
        p = pd[k.split("/", 1)[1]]
        if p.grad is None:
            assert ref.abs().max() == 0, k
            continue
        err = (p.grad.cpu() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-12)
        worst = max(worst, err)
        # table rows are scattered to by cell: rays whose traced point lands in a neighbouring cell move gradient
        # mass between rows, so the hash tables get a looser bound than the dense weights
        bound = 1.0 if "embedding.weight" in k else 0.08
        assert err <= bound + (0.2 if flips else 0.0), (k, err)


@pytest.mark.parametrize("tag", list(IDR_CFGS))
def test_differentiable_part_given_same_trace(golden, tag):
    """Isolates the differentiable part (encode + MLP fwd/bwd + eikonal double backward + loss): the oracle is fed the
    product's own traced distances/masks, so no ray can land in a different hash cell and tolerances are tight:
    outputs abs 5e-5, losses rel 5e-4, parameter gradients 1 % of max-abs (NFFB 3 %)."""
    from idrk.model.implicit_differentiable_renderer import IDRNetwork
    from idrk.model.loss import IDRLoss
    from oracle import idr_oracle as O
    from tests_support import RAY_TRACER_CONF
    g = golden("idr_step")
    et, L, log2T, base, des, bound = IDR_CFGS[tag]
    model = quiet_build(IDRNetwork, make_conf(et, L, log2T, base, des, bound, width=96, feature=32))
    sd = sd_from(g, "sd_%s/" % tag)
    load_sd_into(model, sd)
    model = model.to(DEV).train()
    eik, u = T(g["eik_points_" + tag]), T(g["min_sdf_steps_" + tag])
    model.injected_eikonal_points, model.ray_tracer.injected_min_sdf_steps = eik, u
    inp = {"uv": T(g["uv"]), "pose": T(g["pose"]), "intrinsics": T(g["K"]), "object_mask": T(g["mask"])}
    traced = model.trace({k: v.to(DEV) for k, v in inp.items()})
    out = model.shade(traced)
    lo = IDRLoss(0.1, 100.0, 50.0)(out, {"rgb": T(g["rgb_gt"]).to(DEV)})
    lo["loss"].backward()

    cfg = O.IDRCfg(O.EmbedCfg(et, L, log2T, 2, base, des, bound), ray_tracer=dict(RAY_TRACER_CONF))
    for k, v in sd.items():
        if v.dtype == torch.float32 and not k.endswith(".B"):
            v.requires_grad_(True)
    tr_out = (out["points"].detach().cpu(), traced["network_object_mask"].cpu(), traced["dists"].cpu())
    oout = O.idr_forward(inp, sd, cfg, True, eik, u, tracer_out=tr_out)
    olo = O.idr_loss(oout, T(g["rgb_gt"]))
    nffb = tag == "style"
    assert (out["sdf_output"].cpu() - oout["sdf_output"]).abs().max().item() <= 5e-5
    assert (out["rgb_values"].cpu() - oout["rgb_values"]).abs().max().item() <= (2e-3 if nffb else 2e-4)
    gt_ref = oout["grad_theta"]
    assert (out["grad_theta"].cpu() - gt_ref).abs().max().item() <= (2e-3 if nffb else 2e-4) * gt_ref.abs().max().item()
    for k in ("loss", "rgb_loss", "eikonal_loss", "mask_loss"):
        assert abs(float(lo[k]) - float(olo[k])) <= 5e-4 * max(1.0, abs(float(olo[k]))), k
    names = [k for k, v in sd.items() if v.requires_grad]
    grads = torch.autograd.grad(olo["loss"], [sd[k] for k in names], allow_unused=True)
    pd = dict(model.named_parameters())
    for k, gq in zip(names, grads):
        p = pd[k]
        if gq is None or gq.abs().max() == 0:
            assert p.grad is None or p.grad.abs().max().item() <= 1e-7, k
            continue
        err = (p.grad.cpu() - gq).abs().max().item() / gq.abs().max().item()
        assert err <= (0.03 if nffb else 0.01), (k, err)
