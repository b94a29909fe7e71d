"""Eval-side consumers: eval branch of IDRNetwork.forward, full-image rendering in
splits, SDF grid sweeps for marching cubes.  GPU tests compare with the oracle on the same traced rays
(sdf abs 5e-5, rgb abs 2e-4 / NFFB 2e-3) and check that chunked / split evaluation is bit-identical to the
one-shot call; CPU tests pin the grid construction to an independent numpy restatement of plots.py:226-271."""
import numpy as np
import pytest
import torch

from conftest import sd_from
from tests_support import load_sd_into, make_conf, quiet_build

DEV = "cuda"
IDR_CFGS = {"hash": ("HashGrid", 6, 5, 64, 512, 1.0), "style": ("StyleModNFFB", 6, 5, 16, 512, 0.45)}


def T(a):
    return torch.from_numpy(np.array(a))


# ----------------------------------------------------------------------------------------------
# CPU: grids
# ----------------------------------------------------------------------------------------------
def test_grid_uniform_layout():
    from idrk.utils import plots
    g = plots.get_grid_uniform(7, device=None)
    x = np.linspace(-1.0, 1.0, 7)
    pts = g["grid_points"].numpy()
    assert pts.shape == (343, 3) and g["shortest_axis_length"] == 2.0 and g["shortest_axis_index"] == 0
    # meshgrid('xy') order: point index = (j * nx + i) * nz + k  ->  (x_i, y_j, z_k)
    for (j, i, k) in [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (6, 5, 4), (3, 6, 2)]:
        assert np.allclose(pts[(j * 7 + i) * 7 + k], np.array([x[i], x[j], x[k]], dtype=np.float32))


@pytest.mark.parametrize("short", [0, 1, 2])
def test_grid_around_points(short):
    from idrk.utils import plots
    ext = [1.0, 1.3, 1.7]
    ext[short] = 0.5
    gen = torch.Generator().manual_seed(short)
    p = (torch.rand(500, 3, generator=gen) - 0.5) * torch.tensor(ext)
    g = plots.get_grid(p, 16, device=None)
    lo, hi = p.min(0).values.numpy(), p.max(0).values.numpy()
    assert g["shortest_axis_index"] == short
    ax = g["xyz"][short]
    assert ax.shape[0] == 16 and np.isclose(ax[0], lo[short] - 0.2) and np.isclose(ax[-1], hi[short] + 0.2)
    step = (ax[-1] - ax[0]) / 15
    assert np.isclose(g["shortest_axis_length"], ax[-1] - ax[0])
    for d in range(3):
        a = g["xyz"][d]
        assert np.allclose(np.diff(a), step)
        assert a[0] <= lo[d] - 0.2 + 1e-9 and a[-1] >= hi[d] + 0.2 - 1e-9      # covers the padded box
    n = [len(a) for a in g["xyz"]]
    assert g["grid_points"].shape == (n[0] * n[1] * n[2], 3)


def test_split_merge_roundtrip():
    from idrk.utils.general import merge_output, split_input
    B, N = 2, 2503
    gen = torch.Generator().manual_seed(0)
    inp = {"uv": torch.rand(B, N, 2, generator=gen), "object_mask": torch.rand(B, N, generator=gen) > 0.5,
           "pose": torch.eye(4).repeat(B, 1, 1), "intrinsics": torch.eye(4).repeat(B, 1, 1)}
    parts = split_input(inp, N, n_pixels=700)
    assert [p["uv"].shape[1] for p in parts] == [700, 700, 700, 403]
    res = [{"rgb_values": torch.cat([p["uv"], p["uv"][..., :1]], -1).reshape(-1, 3),
            "mask": p["object_mask"].reshape(-1), "none": None} for p in parts]
    m = merge_output(res, N, B)
    assert "none" not in m
    assert torch.equal(m["rgb_values"].reshape(B, N, 3)[..., :2], inp["uv"])
    assert torch.equal(m["mask"].reshape(B, N), inp["object_mask"])


# ----------------------------------------------------------------------------------------------
# GPU
# ----------------------------------------------------------------------------------------------
def _model(golden, tag):
    from idrk.model.implicit_differentiable_renderer import IDRNetwork
    g = golden("idr_step")
    et, L, log2T, base, des, bound = IDR_CFGS[tag]
    model = quiet_build(IDRNetwork, make_conf(et, L, log2T, base, des, bound, width=96, feature=32))
    sd = sd_from(g, "sd_%s/" % tag)
    load_sd_into(model, sd)
    return model.to(DEV), sd, g


@pytest.mark.gpu
@pytest.mark.parametrize("tag", list(IDR_CFGS))
def test_eval_forward_given_same_trace(golden, tag):
    """IDRNetwork.forward in eval mode (implicit_differentiable_renderer.py:299-302): surface = network mask only,
    no sample network, grad_theta None.  The oracle is fed the product's traced distances / masks."""
    from oracle import idr_oracle as O
    from tests_support import RAY_TRACER_CONF
    model, sd, g = _model(golden, tag)
    model.eval()
    inp = {"uv": T(g["uv"]), "pose": T(g["pose"]), "intrinsics": T(g["K"]), "object_mask": T(g["mask"])}
    traced = model.trace({k: v.to(DEV) for k, v in inp.items()})
    out = model.shade(traced)
    assert out["grad_theta"] is None
    et, L, log2T, base, des, bound = IDR_CFGS[tag]
    cfg = O.IDRCfg(O.EmbedCfg(et, L, log2T, 2, base, des, bound), ray_tracer=dict(RAY_TRACER_CONF))
    tr_out = (out["points"].detach().cpu(), traced["network_object_mask"].cpu(), traced["dists"].cpu())
    oout = O.idr_forward(inp, sd, cfg, False, tracer_out=tr_out)
    nffb = tag == "style"
    assert torch.equal(out["network_object_mask"].cpu(), oout["network_object_mask"])
    assert (out["sdf_output"].detach().cpu() - oout["sdf_output"].detach()).abs().max().item() <= 5e-5
    d = (out["rgb_values"].detach().cpu() - oout["rgb_values"].detach()).abs().max().item()
    assert d <= (2e-3 if nffb else 2e-4), d
    miss = ~traced["network_object_mask"].cpu()
    assert torch.equal(out["rgb_values"].detach().cpu()[miss], torch.ones(int(miss.sum()), 3))
    # eval-mode tracing agrees with the oracle's own eval-mode tracer up to borderline rays
    full = O.idr_forward(inp, sd, cfg, False)
    flips = (full["network_object_mask"] != traced["network_object_mask"].cpu()).sum().item()
    assert flips <= max(1, miss.numel() // 100), flips


@pytest.mark.gpu
def test_render_image_splits_equal_one_shot(golden):
    """general.render_image (eval.py:150-160): 4 splits of <= 70 rays merged == the same rays in one call.
    Masks identical; rgb equal to fp32 rounding of the contraction kernels (tile / split-K choices depend on the row
    count, so bit equality across batch sizes is not promised): abs 2e-5."""
    from idrk.utils.general import render_image
    model, sd, g = _model(golden, "hash")
    inp = {"uv": T(g["uv"]).to(DEV), "pose": T(g["pose"]).to(DEV), "intrinsics": T(g["K"]).to(DEV),
           "object_mask": T(g["mask"]).to(DEV)}
    n = inp["uv"].shape[1]
    model.train()
    merged = render_image(model, inp, n, n_pixels=70, keys=("rgb_values", "network_object_mask", "sdf_output"))
    assert model.training                       # restored
    model.eval()
    one = model(inp)
    assert merged["rgb_values"].shape == (n, 3) and merged["network_object_mask"].shape == (n,)
    flips = (merged["network_object_mask"] != one["network_object_mask"]).sum().item()
    assert flips <= 1, flips
    same = merged["network_object_mask"] == one["network_object_mask"]
    assert (merged["rgb_values"] - one["rgb_values"].detach())[same].abs().max().item() <= 2e-4
    assert (merged["sdf_output"] - one["sdf_output"].detach())[same].abs().max().item() <= 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("tag", list(IDR_CFGS))
def test_sdf_volume_vs_oracle(golden, tag):
    """SDF grid sweep for marching cubes (plots.py:110-130): device-resident chunks through the SDF-only pipeline ==
    the oracle's ImplicitNetwork column 0 (abs 5e-5), chunking is bit-neutral, and the volume has the layout the
    reference passes to measure.marching_cubes."""
    from idrk.utils import plots
    from oracle import idr_oracle as O
    model, sd, g = _model(golden, tag)
    model.eval()
    et, L, log2T, base, des, bound = IDR_CFGS[tag]
    ecfg = O.EmbedCfg(et, L, log2T, 2, base, des, bound)
    res = 20
    grid = plots.get_grid_uniform(res)
    assert grid["grid_points"].is_cuda
    z = plots.sdf_sweep(model.implicit_network, grid["grid_points"], chunk=3000)
    z1 = plots.sdf_sweep(model.implicit_network, grid["grid_points"], chunk=1 << 18)
    assert (z - z1).abs().max().item() <= 2e-6
    with torch.no_grad():
        sub = torch.arange(0, res ** 3, 7)
        ref = O.implicit_forward(grid["grid_points"].cpu()[sub], sd, ecfg, 9, (4,))[:, 0]
    assert (z.cpu()[sub] - ref).abs().max().item() <= 5e-5
    # the reference's callable form gives the same values
    z2 = plots.sdf_sweep(lambda x: model.implicit_network(x)[:, 0], grid["grid_points"], chunk=3000)
    assert (z2 - z).abs().max().item() <= 2e-5
    vol, spacing, origin = plots.sdf_volume(model.implicit_network, grid)
    assert vol is not None and vol.shape == (res, res, res) and vol.dtype == np.float32
    want = z1.cpu().numpy().reshape(res, res, res).transpose([1, 0, 2])
    assert np.array_equal(vol, want)
    x = grid["xyz"][0]
    assert np.allclose(spacing, x[2] - x[1]) and np.allclose(origin, [-1, -1, -1])
    i, j, k = 3, 11, 17
    p = torch.tensor([[x[i], x[j], x[k]]], dtype=torch.float32, device=DEV)
    assert abs(float(model.implicit_network.sdf(p)[0]) - float(vol[i, j, k])) <= 2e-6


# ----------------------------------------------------------------------------------------------
# CPU: the helpers against outputs of the REAL reference functions (tests/golden/eval_helpers.npz)
# ----------------------------------------------------------------------------------------------
def test_grids_equal_reference_functions(golden):
    """utils/plots.py get_grid_uniform / get_grid of the reference (generated by tests/golden/make_golden.py) ==
    idrk.utils.plots, bit for bit: point order, axes, shortest axis and its length."""
    from idrk.utils import plots
    g = golden("eval_helpers")
    assert np.array_equal(plots.get_grid_uniform(5, device=None)["grid_points"].numpy(), g["uniform5_points"])
    for short in range(3):
        o = plots.get_grid(T(g["cloud_%d" % short]), 6, device=None)
        assert np.array_equal(o["grid_points"].numpy(), g["grid_%d_points" % short])
        assert o["shortest_axis_index"] == int(g["grid_%d_meta" % short][0]) == short
        assert float(o["shortest_axis_length"]) == float(g["grid_%d_meta" % short][1])
        for d in range(3):
            assert np.array_equal(np.asarray(o["xyz"][d], dtype=np.float64), g["grid_%d_axis%d" % (short, d)])


def test_split_merge_equal_reference_functions(golden):
    """utils/general.py split_input (10 000-pixel chunks) / merge_output of the reference == ours on the same input."""
    from idrk.utils.general import merge_output, split_input
    g = golden("eval_helpers")
    B, N = 2, 23000
    gen = torch.Generator().manual_seed(4)
    inp = {"uv": torch.rand(B, N, 2, generator=gen), "object_mask": torch.rand(B, N, generator=gen) > 0.5,
           "pose": torch.eye(4).repeat(B, 1, 1), "intrinsics": torch.eye(4).repeat(B, 1, 1)}
    parts = split_input(inp, N)                      # default n_pixels = 10000 like the reference
    assert [p["uv"].shape[1] for p in parts] == list(g["split_sizes"]) == [10000, 10000, 3000]
    res = [{"rgb_values": torch.cat([p["uv"], p["uv"][..., :1]], -1).reshape(-1, 3) * (i + 1),
            "flag": p["object_mask"].reshape(-1).float()} for i, p in enumerate(parts)]
    m = merge_output(res, N, B)
    idx = torch.from_numpy(g["merged_idx"])
    assert np.array_equal(m["rgb_values"][idx].numpy(), g["merged_rgb_rows"])
    assert np.array_equal(m["flag"][idx].numpy(), g["merged_flag_rows"])
    assert np.allclose(m["rgb_values"].double().sum(0).numpy(), g["merged_rgb_sum"], rtol=0, atol=1e-9)
    assert float(m["flag"].double().sum()) == float(g["merged_flag_sum"][0])
