"""Generate the committed golden vectors by running the REAL reference on CPU.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Writes tests/golden/*.npz.  Every fixture holds the inputs, the reference's own weights
(state_dict) and the reference's outputs, so that tests on any box (no reference tree) can
pin oracle/ and the CUDA path against what the reference really computes.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _ref_shim as shim  # noqa: E402


def to_np(sd):
    return {k: v.detach().cpu().numpy() for k, v in sd.items()}


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


def pack_sd(prefix, sd):
    return {prefix + k: v for k, v in to_np(sd).items()}


def gen_hash(R):
    """hash_func / _HashGridMLP / MultiResHashGridMLP  (hashGridEmbedding.py)."""
    g = torch.Generator().manual_seed(11)
    x = torch.rand(257, 3, generator=g) * 2.4 - 1.2          # includes negative coordinates
    x[0] = torch.tensor([0.0, 0.0, 0.0])
    x[1] = torch.tensor([-1e-3, 0.999999, -0.999999])
    out = {"x": x.numpy()}
    # raw hash indices of the 8 corners, a power-of-two table and a res^3 table
    for tag, res, T in (("a", 22, 22 ** 3), ("b", 406, 2 ** 19), ("c", 512, 32)):
        lvl = R.hge._HashGridMLP(3, 2, T, res)
        xi = (x * res).long().unsqueeze(-2)
        bm = lvl.bin_mask.reshape((1,) + lvl.bin_mask.shape)
        inds = torch.where(bm, xi, xi + 1)
        out["idx_" + tag] = R.hge.hash_func(inds, lvl.primes, T).numpy()
        out["meta_" + tag] = np.array([res, T], dtype=np.int64)
    # level schedules
    for tag, args in (("s1", (True, 3, 16, 2, 19, 16, 2048)), ("s2", (True, 3, 6, 2, 5, 64, 512)),
                      ("s3", (True, 3, 4, 2, 3, 16, 512)), ("s4", (True, 3, 6, 2, 22, 16, 512))):
        if tag in ("s1", "s4"):
            # do not allocate the big tables: recompute the schedule exactly as the ctor does
            import math
            _, _, L, _, log2T, base, desired = args
            beta = math.exp((math.log(desired) - math.log(base)) / (L - 1))
            res = [math.floor(base * (beta ** l)) for l in range(L)]
            rows = [min(r ** 3, 2 ** log2T) for r in res]
        else:
            with shim.quiet():
                m = R.hge.MultiResHashGridMLP(*args)
            res = [int(l.resolution) for l in m.levels]
            rows = [int(l.hashmap_size) for l in m.levels]
        out["sched_" + tag] = np.array([res, rows], dtype=np.int64)
        out["sched_args_" + tag] = np.array(args[2:], dtype=np.int64)
    # full embeddings, two configs, with the reference's own weights
    for tag, args in (("e1", (True, 3, 16, 2, 10, 16, 2048)), ("e2", (True, 3, 6, 2, 5, 64, 512))):
        torch.manual_seed(5)
        with shim.quiet():
            m = R.hge.MultiResHashGridMLP(*args)
        y = m(x)
        out["emb_" + tag] = y.detach().numpy()
        out["emb_args_" + tag] = np.array(args[2:], dtype=np.int64)
        out.update(pack_sd("sd_%s/" % tag, m.state_dict()))
        # table gradient of sum(y * w) with fixed w
        w = torch.rand(y.shape, generator=torch.Generator().manual_seed(3))
        (y * w).sum().backward()
        out["w_" + tag] = w.numpy()
        for l, lvl in enumerate(m.levels):
            out["grad_%s/%d" % (tag, l)] = lvl.embedding.weight.grad.numpy()
    save("hashgrid", **out)


def gen_encoders(R):
    """PositionalEncoding, FourierFeature, FourierFilterBanks (FFB + StyleModNFFB), selector."""
    g = torch.Generator().manual_seed(21)
    x = (torch.rand(193, 3, generator=g) * 2 - 1) * 0.44
    out = {"x": x.numpy()}
    pe = R.fe.PositionalEncoding(include_input=True, input_dims=3, max_freq_log2=5, num_freqs=6,
                                 log_sampling=True, periodic_fns=[torch.sin, torch.cos])
    out["posenc_6_5"] = pe(x).numpy()
    emb_fn, od = R.fe.get_embedder(4)
    out["view_nerfpos4"] = emb_fn(x).numpy()
    out["view_nerfpos4_outdim"] = np.array([od])
    for tag, etype, L in (("ffb", "FFB", 6), ("style", "StyleModNFFB", 6), ("ffb4", "FFB", 4)):
        torch.manual_seed(7)
        with shim.quiet():
            net = R.ced.Custom_Embedding_Network(3, [3, 64], etype, L, 5 if L == 6 else 3, 2, 16, 512, 0.45 if L == 6 else 1.0)
        # give the out_layer / tables non-trivial values
        with torch.no_grad():
            for n_, p in net.named_parameters():
                if "embedding.weight" in n_:
                    p.mul_(3000.0)
        xr = x.clone().requires_grad_(True)
        y = net(xr)
        out["emb_" + tag] = y.detach().numpy()
        out["dim_" + tag] = np.array([net.embeddings_dim])
        w = torch.rand(y.shape, generator=torch.Generator().manual_seed(4))
        gx = torch.autograd.grad((y * w).sum(), xr, retain_graph=True)[0]
        out["w_" + tag] = w.numpy()
        out["gx_" + tag] = gx.numpy()
        out.update(pack_sd("sd_%s/" % tag, net.state_dict()))
    save("encoders", **out)


def small_conf(R, embed_type, multires, log2T, base, desired, bound, view_type="NerfPos", width=96):
    return R.DictConf({
        "feature_vector_size": 32,
        "implicit_network": {"d_in": 3, "d_out": 1, "dims": [width] * 8, "geometric_init": True, "bias": 0.6,
                             "skip_in": [4], "weight_norm": True, "multires": multires},
        "rendering_network": {"mode": "idr", "d_in": 9, "d_out": 3, "viewdirs_embed_type": view_type,
                              "dims": [width] * 4, "weight_norm": True, "multires_view": 4},
        "ray_tracer": {"object_bounding_sphere": 1.0, "sdf_threshold": 5.0e-5, "line_search_step": 0.5,
                       "line_step_iters": 3, "sphere_tracing_iters": 10, "n_steps": 100, "n_secant_steps": 8},
        "embedding_network": {"embed_type": embed_type, "log2_max_hash_size": log2T, "max_points_per_entry": 2,
                              "base_resolution": base, "desired_resolution": desired, "bound": bound},
    })


def perturb_model(model, seed, scale=0.02):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n_, p in model.named_parameters():
            if "embedding.weight" in n_:
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * 0.3)
            elif p.dim() >= 1 and "weight_g" not in n_:
                p.add_(torch.randn(p.shape, generator=g) * scale)


def gen_networks(R):
    """ImplicitNetwork.forward / .gradient (+ eikonal double backward), RenderingNetwork."""
    out = {}
    g = torch.Generator().manual_seed(31)
    x = (torch.rand(129, 3, generator=g) * 2 - 1) * 0.9
    out["x"] = x.numpy()
    for tag, (et, L, log2T, base, des, bound) in {
            "hash": ("HashGrid", 6, 5, 64, 512, 1.0),
            "hash16": ("HashGrid", 16, 8, 16, 2048, 1.0),
            "ffb": ("FFB", 6, 5, 16, 512, 0.45),
            "style": ("StyleModNFFB", 6, 5, 16, 512, 0.45)}.items():
        conf = small_conf(R, et, L, log2T, base, des, bound)
        torch.manual_seed(3)
        with shim.quiet():
            model = R.idr.IDRNetwork(conf)
        perturb_model(model, 17)
        net = model.implicit_network
        y = net(x)
        out["y_" + tag] = y.detach().numpy()
        xg = x.clone()
        gr = net.gradient(xg)
        out["grad_" + tag] = gr.detach().numpy()
        # eikonal-style loss on the gradient -> parameter gradients (double backward)
        loss = ((gr[:, 0, :].norm(2, dim=1) - 1) ** 2).mean() + y[:, 0].mean() + 0.01 * (y[:, 1:] ** 2).mean()
        params = [(n_, p) for n_, p in net.named_parameters() if p.requires_grad]
        grads = torch.autograd.grad(loss, [p for _, p in params], allow_unused=True)
        out["loss_" + tag] = np.array([loss.item()])
        for (n_, _), gq in zip(params, grads):
            if gq is not None and gq.numel() <= 70000:
                out["pg_%s/%s" % (tag, n_)] = gq.numpy()
        out.update(pack_sd("sd_%s/" % tag, model.state_dict()))
        # rendering network
        gg = torch.Generator().manual_seed(5)
        pts = torch.rand(65, 3, generator=gg) - 0.5
        nrm = torch.nn.functional.normalize(torch.randn(65, 3, generator=gg), dim=1)
        vd = torch.nn.functional.normalize(torch.randn(65, 3, generator=gg), dim=1)
        ft = torch.randn(65, 32, generator=gg)
        rgb = model.rendering_network(pts, nrm, vd, ft)
        out["rn_in_" + tag] = np.concatenate([pts.numpy(), nrm.numpy(), vd.numpy(), ft.numpy()], 1)
        out["rn_out_" + tag] = rgb.detach().numpy()
    # deep view embedder variant
    conf = small_conf(R, "FFB", 6, 5, 16, 512, 0.45, view_type="FFB")
    torch.manual_seed(4)
    with shim.quiet():
        model = R.idr.IDRNetwork(conf)
    perturb_model(model, 19)
    gg = torch.Generator().manual_seed(6)
    pts = torch.rand(65, 3, generator=gg) - 0.5
    nrm = torch.nn.functional.normalize(torch.randn(65, 3, generator=gg), dim=1)
    vd = torch.nn.functional.normalize(torch.randn(65, 3, generator=gg), dim=1)
    ft = torch.randn(65, 32, generator=gg)
    out["rn_in_viewffb"] = np.concatenate([pts.numpy(), nrm.numpy(), vd.numpy(), ft.numpy()], 1)
    out["rn_out_viewffb"] = model.rendering_network(pts, nrm, vd, ft).detach().numpy()
    out.update(pack_sd("sd_viewffb/", model.state_dict()))
    save("networks", **out)


def analytic_sdf(kind):
    """Analytic SDFs evaluated with plain fp32 torch ops (both sides use the same function)."""
    if kind == "sphere":
        return lambda p: p.norm(2, dim=1) - 0.5
    if kind == "bumpy":
        return lambda p: (p.norm(2, dim=1) - 0.55) * 0.7 + 0.05 * torch.sin(9.0 * p[:, 0]) * torch.sin(7.0 * p[:, 1])
    if kind == "torus":
        def f(p):
            q = torch.stack([torch.sqrt(p[:, 0] ** 2 + p[:, 2] ** 2) - 0.45, p[:, 1]], 1)
            return q.norm(2, dim=1) - 0.18
        return f
    raise KeyError(kind)


def synth_rays(n, seed, tz=-3.0):
    pose = torch.eye(4).unsqueeze(0)
    pose[0, 2, 3] = tz
    K = torch.eye(4).unsqueeze(0)
    K[0, 0, 0] = K[0, 1, 1] = 500.0
    K[0, 0, 2] = K[0, 1, 2] = 128.0
    uv = torch.rand(1, n, 2, generator=torch.Generator().manual_seed(seed)) * 256
    mask = torch.rand(1, n, generator=torch.Generator().manual_seed(seed + 1)) > 0.5
    return uv, pose, K, mask


def gen_raytracing(R):
    """get_camera_params, get_sphere_intersection, RayTracing.forward with analytic SDFs."""
    out = {}
    uv, pose, K, mask = synth_rays(600, 1)
    dirs, cam = R.ru.get_camera_params(uv, pose, K)
    out["uv"], out["pose"], out["K"], out["mask"] = uv.numpy(), pose.numpy(), K.numpy(), mask.numpy()
    out["dirs"], out["cam"] = dirs.numpy(), cam.numpy()
    t, hit = R.ru.get_sphere_intersection(cam, dirs, r=1.0)
    out["sph_t"], out["sph_hit"] = t.numpy(), hit.numpy()
    for kind in ("sphere", "bumpy", "torus"):
        for training in (True, False):
            tr = R.rt.RayTracing(1.0, 5.0e-5, 0.5, 3, 10, 100, 8)
            tr.train(training)
            torch.manual_seed(77)
            with shim.quiet(), torch.no_grad():
                pts, nm, d = tr(sdf=analytic_sdf(kind), cam_loc=cam, object_mask=mask.reshape(-1), ray_directions=dirs)
            tag = "%s_%s" % (kind, "train" if training else "eval")
            out["pts_" + tag], out["net_" + tag], out["dist_" + tag] = pts.numpy(), nm.numpy(), d.numpy()
    torch.manual_seed(77)
    out["min_sdf_steps"] = torch.empty(100).uniform_(0.0, 1.0).numpy()
    save("raytracing", **out)


def gen_idr(R):
    """IDRNetwork.forward + IDRLoss + backward on a small model (width 96) with 256 rays."""
    out = {}
    uv, pose, K, mask = synth_rays(256, 1)
    rgb_gt = torch.rand(1, 256, 3, generator=torch.Generator().manual_seed(3)) * 2 - 1
    out["uv"], out["pose"], out["K"], out["mask"], out["rgb_gt"] = (uv.numpy(), pose.numpy(), K.numpy(),
                                                                    mask.numpy(), rgb_gt.numpy())
    for tag, (et, L, log2T, base, des, bound) in {
            "hash": ("HashGrid", 6, 5, 64, 512, 1.0),
            "style": ("StyleModNFFB", 6, 5, 16, 512, 0.45)}.items():
        conf = small_conf(R, et, L, log2T, base, des, bound)
        torch.manual_seed(3)
        with shim.quiet():
            model = R.idr.IDRNetwork(conf)
        perturb_model(model, 23, scale=0.01)
        loss_fn = R.loss.IDRLoss(eikonal_weight=0.1, mask_weight=100.0, alpha=50.0)
        model.train()
        torch.manual_seed(99)
        with shim.quiet():
            o = model({"uv": uv, "pose": pose, "intrinsics": K, "object_mask": mask})
            lo = loss_fn(o, {"rgb": rgb_gt})
        lo["loss"].backward()
        # the two CPU-RNG draws, in the reference's call order (ray_tracing.py:277 then idr:279)
        torch.manual_seed(99)
        out["min_sdf_steps_" + tag] = torch.empty(100).uniform_(0.0, 1.0).numpy()
        out["eik_points_" + tag] = torch.empty(128, 3).uniform_(-1.0, 1.0).numpy()
        for k in ("points", "rgb_values", "sdf_output", "network_object_mask", "grad_theta"):
            out["%s_%s" % (k, tag)] = o[k].detach().numpy()
        for k in ("loss", "rgb_loss", "eikonal_loss", "mask_loss"):
            out["%s_%s" % (k, tag)] = np.array([float(lo[k])])
        for n_, p in model.named_parameters():
            if p.grad is not None and p.numel() <= 70000:
                out["pg_%s/%s" % (tag, n_)] = p.grad.numpy()
        out.update(pack_sd("sd_%s/" % tag, model.state_dict()))
    save("idr_step", **out)


def gen_idr_eval(R):
    """IDRNetwork.forward in EVAL mode (implicit_differentiable_renderer.py:299-302; ray_tracing.py:66-69,233 eval
    branches) on the models of idr_step.npz (their state dicts are read back from that fixture, not stored twice)."""
    g = np.load(os.path.join(HERE, "idr_step.npz"))
    uv, pose, K, mask = (torch.from_numpy(g[k]) for k in ("uv", "pose", "K", "mask"))
    out = {}
    for tag, (et, L, log2T, base, des, bound) in {
            "hash": ("HashGrid", 6, 5, 64, 512, 1.0),
            "style": ("StyleModNFFB", 6, 5, 16, 512, 0.45)}.items():
        conf = small_conf(R, et, L, log2T, base, des, bound)
        with shim.quiet():
            model = R.idr.IDRNetwork(conf)
        sd = {k[len("sd_%s/" % tag):]: torch.from_numpy(np.array(g[k])) for k in g.files if k.startswith("sd_%s/" % tag)}
        model.load_state_dict(sd)
        model.eval()
        with shim.quiet():
            o = model({"uv": uv, "pose": pose, "intrinsics": K, "object_mask": mask})
        for k in ("points", "rgb_values", "sdf_output", "network_object_mask"):
            out["%s_%s" % (k, tag)] = o[k].detach().numpy()
        assert o["grad_theta"] is None
    save("idr_eval", **out)


def _import_ref_utils():
    """utils.plots / utils.general of the reference with its absent plotting / mesh dependencies stubbed (only the
    grid and split helpers are called)."""
    import types
    for name in ("plotly", "plotly.graph_objs", "plotly.offline", "trimesh", "torchvision", "PIL", "PIL.Image"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    sys.modules["plotly"].graph_objs = sys.modules["plotly.graph_objs"]
    sys.modules["plotly"].offline = sys.modules["plotly.offline"]
    sk = sys.modules["skimage"]
    sk.measure = types.ModuleType("skimage.measure")
    sys.modules["skimage.measure"] = sk.measure
    if not hasattr(sys.modules["PIL"], "Image"):
        sys.modules["PIL"].Image = sys.modules["PIL.Image"]
    import utils.general as rg
    import utils.plots as rp
    return rp, rg


def gen_eval_helpers(R):
    """Grids of utils/plots.py:226-271 and the split / merge helpers of utils/general.py:23-50, from the reference."""
    rp, rg = _import_ref_utils()
    out = {}
    gu = rp.get_grid_uniform(5)
    out["uniform5_points"] = gu["grid_points"].numpy()
    for short in range(3):
        ext = [1.0, 1.3, 1.7]
        ext[short] = 0.5
        pts = (torch.rand(200, 3, generator=torch.Generator().manual_seed(short)) - 0.5) * torch.tensor(ext)
        gg = rp.get_grid(pts, 6)
        out["cloud_%d" % short] = pts.numpy()
        out["grid_%d_points" % short] = gg["grid_points"].numpy()
        out["grid_%d_meta" % short] = np.array([gg["shortest_axis_index"], gg["shortest_axis_length"]], dtype=np.float64)
        for d in range(3):
            out["grid_%d_axis%d" % (short, d)] = np.asarray(gg["xyz"][d], dtype=np.float64)
    B, N = 2, 23000                                        # the reference splits in chunks of 10 000 pixels
    gen = torch.Generator().manual_seed(4)
    inp = {"uv": torch.rand(B, N, 2, generator=gen), "object_mask": torch.rand(B, N, generator=gen) > 0.5,
           "pose": torch.eye(4).repeat(B, 1, 1), "intrinsics": torch.eye(4).repeat(B, 1, 1)}
    parts = rg.split_input(inp, N)
    out["split_sizes"] = np.array([p["uv"].shape[1] for p in parts])
    res = [{"rgb_values": torch.cat([p["uv"], p["uv"][..., :1]], -1).reshape(-1, 3) * (i + 1),
            "flag": p["object_mask"].reshape(-1).float()} for i, p in enumerate(parts)]
    merged = rg.merge_output(res, N, B)
    # inputs are regenerated from the seed by the test; of the merged [B * N, .] outputs keep rows around the split
    # boundaries of both batch entries and the column sums (small fixture)
    idx = np.array([b * N + i for b in range(B) for i in (0, 1, 9999, 10000, 10001, 19999, 20000, 22999)])
    out["merged_idx"] = idx
    out["merged_rgb_rows"], out["merged_flag_rows"] = merged["rgb_values"].numpy()[idx], merged["flag"].numpy()[idx]
    out["merged_rgb_sum"] = merged["rgb_values"].double().sum(0).numpy()
    out["merged_flag_sum"] = np.array([merged["flag"].double().sum().item()])
    save("eval_helpers", **out)


def main():
    if len(sys.argv) > 1 and sys.argv[1] in ("idr_eval", "eval_helpers"):   # add one fixture without rewriting the others
        if not shim.available():
            raise SystemExit("reference tree not found")
        torch.set_num_threads(8)
        {"idr_eval": gen_idr_eval, "eval_helpers": gen_eval_helpers}[sys.argv[1]](shim.load())
        return
    if not shim.available():
        raise SystemExit("reference tree not found; golden vectors can only be generated in the build container")
    torch.set_num_threads(8)
    R = shim.load()
    gen_hash(R)
    gen_encoders(R)
    gen_networks(R)
    gen_raytracing(R)
    gen_idr(R)
    gen_idr_eval(R)
    gen_eval_helpers(R)


if __name__ == "__main__":
    main()
