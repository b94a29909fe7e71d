"""Pins oracle/ against the golden vectors produced by the real reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import sd_from
from oracle import idr_oracle as O


def T(a):
    return torch.from_numpy(np.array(a))


def test_level_schedule(golden):
    g = golden("hashgrid")
    for tag in ("s1", "s2", "s3", "s4"):
        L, F, log2T, base, desired = [int(v) for v in g["sched_args_" + tag]]
        res, rows = O.level_schedule(L, base, desired, log2T)
        assert res == g["sched_" + tag][0].tolist()
        assert rows == g["sched_" + tag][1].tolist()


def test_hash_indices_bit_exact(golden):
    g = golden("hashgrid")
    x = T(g["x"])
    for tag in ("a", "b", "c"):
        res, rows = [int(v) for v in g["meta_" + tag]]
        idx = O.corner_hash(x, res, rows)
        assert np.array_equal(idx, g["idx_" + tag])


@pytest.mark.parametrize("tag", ["e1", "e2"])
def test_hashgrid_embedding_and_table_grads(golden, tag):
    g = golden("hashgrid")
    x = T(g["x"])
    L, F, log2T, base, desired = [int(v) for v in g["emb_args_" + tag]]
    sd = sd_from(g, "sd_%s/" % tag)
    for v in sd.values():
        v.requires_grad_(v.dtype == torch.float32 and v.dim() == 2 and v.shape[1] == F)
    y = O.hashgrid_embed(x, sd, "", L, base, desired)
    ref = T(g["emb_" + tag])
    assert y.shape == ref.shape
    # hash block: bit exact.  Fourier block: same torch ops -> exact on the same build
    assert torch.equal(y[:, 3 + 2 * L:], ref[:, 3 + 2 * L:])
    assert torch.allclose(y[:, :3 + 2 * L], ref[:, :3 + 2 * L], atol=2e-6, rtol=0)
    (y * T(g["w_" + tag])).sum().backward()
    for l in range(L):
        got = sd["levels.%d.embedding.weight" % l].grad
        assert torch.allclose(got, T(g["grad_%s/%d" % (tag, l)]), atol=1e-6, rtol=1e-5)


def test_posenc(golden):
    g = golden("encoders")
    x = T(g["x"])
    assert torch.allclose(O.positional_encoding(x, 6, 5, True), T(g["posenc_6_5"]), atol=1e-6)
    assert torch.allclose(O.view_embed_nerfpos(x, 4), T(g["view_nerfpos4"]), atol=1e-6)


@pytest.mark.parametrize("tag,etype,L,log2T,bound", [("ffb", "FFB", 6, 5, 0.45), ("style", "StyleModNFFB", 6, 5, 0.45),
                                                     ("ffb4", "FFB", 4, 3, 1.0)])
def test_nffb(golden, tag, etype, L, log2T, bound):
    g = golden("encoders")
    x = T(g["x"]).clone().requires_grad_(True)
    sd = sd_from(g, "sd_%s/" % tag)
    cfg = O.EmbedCfg(etype, L, log2T, 2, 16, 512, bound)
    y = O.embed(x, sd, "", cfg)
    assert y.shape[1] == int(g["dim_" + tag][0]) == cfg.width()
    assert torch.allclose(y, T(g["emb_" + tag]), atol=2e-5, rtol=1e-5)
    gx = torch.autograd.grad((y * T(g["w_" + tag])).sum(), x)[0]
    ref = T(g["gx_" + tag])
    assert torch.allclose(gx, ref, atol=1e-4 * ref.abs().max().item(), rtol=1e-4)


NET_CFGS = {"hash": ("HashGrid", 6, 5, 64, 512, 1.0), "hash16": ("HashGrid", 16, 8, 16, 2048, 1.0),
            "ffb": ("FFB", 6, 5, 16, 512, 0.45), "style": ("StyleModNFFB", 6, 5, 16, 512, 0.45)}


@pytest.mark.parametrize("tag", list(NET_CFGS))
def test_implicit_network(golden, tag):
    g = golden("networks")
    et, L, log2T, base, des, bound = NET_CFGS[tag]
    cfg = O.EmbedCfg(et, L, log2T, 2, base, des, bound)
    sd = sd_from(g, "sd_%s/" % tag)
    for k, v in sd.items():
        if v.dtype == torch.float32 and k.startswith("implicit_network") and not k.endswith(".B"):
            v.requires_grad_(True)
    x = T(g["x"])
    y = O.implicit_forward(x, sd, cfg)
    assert torch.allclose(y, T(g["y_" + tag]), atol=1e-5, rtol=1e-5)
    gr = O.implicit_gradient(x.clone(), sd, cfg)
    ref = T(g["grad_" + tag])
    assert torch.allclose(gr, ref, atol=1e-4 * ref.abs().max().item(), rtol=1e-4)
    loss = ((gr[:, 0, :].norm(2, dim=1) - 1) ** 2).mean() + y[:, 0].mean() + 0.01 * (y[:, 1:] ** 2).mean()
    assert abs(loss.item() - float(g["loss_" + tag][0])) <= 1e-4 * max(1.0, abs(float(g["loss_" + tag][0])))
    names = [k for k in g if k.startswith("pg_%s/" % tag)]
    assert names
    params = [sd["implicit_network." + k.split("/", 1)[1]] for k in names]
    grads = torch.autograd.grad(loss, params, allow_unused=True)
    for k, gq in zip(names, grads):
        ref = T(g[k])
        if gq is None:
            assert ref.abs().max() == 0
            continue
        assert torch.allclose(gq, ref, atol=2e-4 * max(ref.abs().max().item(), 1e-8), rtol=1e-3), k


@pytest.mark.parametrize("tag", list(NET_CFGS) + ["viewffb"])
def test_rendering_network(golden, tag):
    g = golden("networks")
    sd = sd_from(g, "sd_%s/" % tag)
    inp = T(g["rn_in_" + tag])
    view_cfg = O.EmbedCfg("FFB", 4, 3, 2, 16, 512, 1.0) if tag == "viewffb" else None
    rgb = O.rendering_forward(inp[:, 0:3], inp[:, 3:6], inp[:, 6:9], inp[:, 9:], sd, 4, 5, view_cfg)
    assert torch.allclose(rgb, T(g["rn_out_" + tag]), atol=2e-5, rtol=1e-5)


def test_camera_and_sphere(golden):
    g = golden("raytracing")
    dirs, cam = O.camera_rays(T(g["uv"]), T(g["pose"]), T(g["K"]))
    assert torch.allclose(dirs, T(g["dirs"]), atol=1e-7)
    assert torch.equal(cam, T(g["cam"]))
    t, hit = O.sphere_intersection(cam, dirs, 1.0)
    assert torch.equal(hit, T(g["sph_hit"]))
    assert torch.allclose(t, T(g["sph_t"]), atol=1e-6)


def _analytic(kind):
    if kind == "sphere":
        return lambda p: p.norm(2, dim=1) - 0.5
    if kind == "bumpy":
        return lambda p: (p.norm(2, dim=1) - 0.55) * 0.7 + 0.05 * torch.sin(9.0 * p[:, 0]) * torch.sin(7.0 * p[:, 1])

    def torus(p):
        q = torch.stack([torch.sqrt(p[:, 0] ** 2 + p[:, 2] ** 2) - 0.45, p[:, 1]], 1)
        return q.norm(2, dim=1) - 0.18
    return torus


@pytest.mark.parametrize("kind", ["sphere", "bumpy", "torus"])
@pytest.mark.parametrize("training", [True, False])
def test_raytracer_bit_exact(golden, kind, training):
    g = golden("raytracing")
    tr = O.RayTracerOracle(1.0, 5.0e-5, 0.5, 3, 10, 100, 8)
    tr.training = training
    # feed the reference's own ray directions so that masks are bit-exact
    pts, nm, d = tr(_analytic(kind), T(g["cam"]), T(g["mask"]).reshape(-1), T(g["dirs"]), T(g["min_sdf_steps"]))
    tag = "%s_%s" % (kind, "train" if training else "eval")
    assert torch.equal(nm, T(g["net_" + tag]))
    assert torch.equal(d, T(g["dist_" + tag]))
    assert torch.equal(pts, T(g["pts_" + tag]))


IDR_CFGS = {"hash": ("HashGrid", 6, 5, 64, 512, 1.0), "style": ("StyleModNFFB", 6, 5, 16, 512, 0.45)}


@pytest.mark.parametrize("tag", list(IDR_CFGS))
def test_idr_step(golden, tag):
    g = golden("idr_step")
    et, L, log2T, base, des, bound = IDR_CFGS[tag]
    cfg = O.IDRCfg(O.EmbedCfg(et, L, log2T, 2, base, des, bound),
                   ray_tracer=dict(object_bounding_sphere=1.0, sdf_threshold=5.0e-5, line_search_step=0.5,
                                   line_step_iters=3, sphere_tracing_iters=10, n_steps=100, n_secant_steps=8))
    sd = sd_from(g, "sd_%s/" % tag)
    for k, v in sd.items():
        if v.dtype == torch.float32 and not k.endswith(".B"):
            v.requires_grad_(True)
    inp = {"uv": T(g["uv"]), "pose": T(g["pose"]), "intrinsics": T(g["K"]), "object_mask": T(g["mask"])}
    out = O.idr_forward(inp, sd, cfg, True, T(g["eik_points_" + tag]), T(g["min_sdf_steps_" + tag]))
    assert torch.equal(out["network_object_mask"], T(g["network_object_mask_" + tag]))
    assert torch.allclose(out["points"], T(g["points_" + tag]), atol=1e-5)
    assert torch.allclose(out["sdf_output"], T(g["sdf_output_" + tag]), atol=1e-5)
    assert torch.allclose(out["rgb_values"], T(g["rgb_values_" + tag]), atol=1e-4)
    gt_ref = T(g["grad_theta_" + tag])
    assert torch.allclose(out["grad_theta"], gt_ref, atol=3e-4 * gt_ref.abs().max().item(), rtol=1e-3)
    lo = O.idr_loss(out, T(g["rgb_gt"]))
    for k in ("loss", "rgb_loss", "eikonal_loss", "mask_loss"):
        assert abs(float(lo[k]) - float(g["%s_%s" % (k, tag)][0])) <= 1e-4 * max(1.0, abs(float(g["%s_%s" % (k, tag)][0]))), k
    names = [k for k in g if k.startswith("pg_%s/" % tag)]
    params = [sd[k.split("/", 1)[1]] for k in names]
    grads = torch.autograd.grad(lo["loss"], params, allow_unused=True)
    for k, gq in zip(names, grads):
        ref = T(g[k])
        if gq is None:
            assert ref.abs().max() == 0, k
            continue
        assert torch.allclose(gq, ref, atol=5e-4 * max(ref.abs().max().item(), 1e-8), rtol=1e-2), k


@pytest.mark.parametrize("tag", list(IDR_CFGS))
def test_idr_eval_forward(golden, tag):
    """Eval-mode forward (implicit_differentiable_renderer.py:299-302; tracer eval branches ray_tracing.py:66-69,233) of
    the oracle vs the real reference on the idr_step models: hit masks bit-exact, points 1e-5, sdf 1e-5, rgb 1e-4."""
    g, e = golden("idr_step"), golden("idr_eval")
    et, L, log2T, base, des, bound = IDR_CFGS[tag]
    cfg = O.IDRCfg(O.EmbedCfg(et, L, log2T, 2, base, des, bound),
                   ray_tracer=dict(object_bounding_sphere=1.0, sdf_threshold=5.0e-5, line_search_step=0.5,
                                   line_step_iters=3, sphere_tracing_iters=10, n_steps=100, n_secant_steps=8))
    sd = sd_from(g, "sd_%s/" % tag)
    inp = {"uv": T(g["uv"]), "pose": T(g["pose"]), "intrinsics": T(g["K"]), "object_mask": T(g["mask"])}
    out = O.idr_forward(inp, sd, cfg, False)
    assert out["grad_theta"] is None
    assert torch.equal(out["network_object_mask"], T(e["network_object_mask_" + tag]))
    assert torch.allclose(out["points"], T(e["points_" + tag]), atol=1e-5)
    assert torch.allclose(out["sdf_output"].detach(), T(e["sdf_output_" + tag]), atol=1e-5)
    assert torch.allclose(out["rgb_values"].detach(), T(e["rgb_values_" + tag]), atol=1e-4)
    # eval mode differs from training mode on this batch (the sampler does not consult the ground-truth mask)
    assert not torch.equal(T(e["network_object_mask_" + tag]), T(g["network_object_mask_" + tag])) or \
        not torch.allclose(T(e["rgb_values_" + tag]), T(g["rgb_values_" + tag]))
