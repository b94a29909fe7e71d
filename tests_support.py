"""Helpers shared by tests/, __graft_entry__.smoke() and bench.py (checker side only)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def synthetic_batch(n_rays: int, seed: int = 1):
    """SURVEY.md section 8(d) synthetic camera and batch: pose = I, t = (0, 0, -3), f = 500, c = 128, uv in [0, 256),
    random object mask and target colours - an INPUT recipe (no reference arithmetic), shared by bench.py, scripts/
    and the oracle-side tests so that every arm sees the same rays."""
    pose = torch.eye(4).unsqueeze(0)
    pose[0, 2, 3] = -3.0
    K = torch.eye(4).unsqueeze(0)
    K[0, 0, 0] = K[0, 1, 1] = 500.0
    K[0, 0, 2] = K[0, 1, 2] = 128.0
    uv = torch.rand(1, n_rays, 2, generator=torch.Generator().manual_seed(seed)) * 256
    mask = torch.rand(1, n_rays, generator=torch.Generator().manual_seed(seed + 1)) > 0.5
    rgb = torch.rand(1, n_rays, 3, generator=torch.Generator().manual_seed(seed + 2)) * 2 - 1
    return {"uv": uv, "pose": pose, "intrinsics": K, "object_mask": mask}, rgb


def load_sd_into(module: torch.nn.Module, sd: dict, prefix: str = "", strict: bool = True):
    """Copies a reference-keyed state dict (optionally under `prefix`) into a product module."""
    sub = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    missing, unexpected = module.load_state_dict(sub, strict=False)
    if strict and (missing or unexpected):
        raise AssertionError("state_dict mismatch: missing=%s unexpected=%s" % (missing, unexpected))
    return module


def smoke_check(device):
    """One small invocation of the hot path on `device`, checked against the oracle: hash encode (bit exact)
    and one IDR training step (ray trace + encode + MLP fwd/bwd + eikonal) on 256 rays."""
    import numpy as np
    from oracle import idr_oracle as O
    from idrk.model.embeddings.hashGridEmbedding import MultiResHashGridMLP
    from idrk.model.implicit_differentiable_renderer import IDRNetwork
    from idrk.model.loss import IDRLoss
    gen = torch.Generator().manual_seed(0)
    L, F, log2T, base, desired = 16, 2, 12, 16, 2048
    sd = O.make_hashgrid_sd("", L, F, log2T, base, desired, gen, table_std=0.5)
    m = MultiResHashGridMLP(True, 3, L, F, log2T, base, desired)
    load_sd_into(m, sd)
    m = m.to(device)
    x = torch.rand(4096, 3, generator=gen) * 2 - 1
    y = m(x.to(device)).cpu()
    ref = O.hashgrid_embed(x, sd, "", L, base, desired)
    assert torch.equal(y[:, 3 + 2 * L:], ref[:, 3 + 2 * L:]), "hash block must be bit exact"
    assert torch.allclose(y[:, :3 + 2 * L], ref[:, :3 + 2 * L], atol=4e-6, rtol=0), "fourier prefix"

    cfg = O.IDRCfg(O.EmbedCfg("HashGrid", 6, 5, 2, 64, 512, 1.0), ray_tracer=dict(RAY_TRACER_CONF))
    conf = make_conf("HashGrid", 6, 5, 64, 512, 1.0, width=128, feature=32)
    model = quiet_build(IDRNetwork, conf)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.to(device).train()
    inp, rgb = O.synthetic_batch(256, seed=1)
    eik = torch.rand(128, 3, generator=gen) * 2 - 1
    u = torch.rand(100, generator=gen)
    model.injected_eikonal_points, model.ray_tracer.injected_min_sdf_steps = eik, u
    out = model({k: v.to(device) for k, v in inp.items()})
    lo = IDRLoss(0.1, 100.0, 50.0)(out, {"rgb": rgb.to(device)})
    lo["loss"].backward()
    for v in sd.values():
        if v.dtype == torch.float32 and v.dim() > 0:
            v.requires_grad_(True)
    oout = O.idr_forward(inp, sd, cfg, True, eik, u)
    olo = O.idr_loss(oout, rgb)
    flips = (out["network_object_mask"].cpu() != oout["network_object_mask"]).sum().item()
    assert flips <= 3, "hit/miss masks: %d flips" % flips
    assert abs(float(lo["loss"].detach()) - float(olo["loss"].detach())) <= 2e-2 * max(1.0, abs(float(olo["loss"].detach())))
    assert model.implicit_network.lin0.weight_v.grad is not None
    return True


class DictConf(dict):
    """Stand-in for a pyhocon ConfigTree (the getters IDRNetwork uses)."""

    def _walk(self, key):
        node = self
        for part in key.split("."):
            node = node[part]
        return node

    def get_int(self, key):
        return int(self._walk(key))

    def get_float(self, key):
        return float(self._walk(key))

    def get_string(self, key):
        return str(self._walk(key))

    def get_list(self, key):
        return list(self._walk(key))

    def get_config(self, key):
        try:
            node = self._walk(key)
        except KeyError:
            return None
        return DictConf(node) if isinstance(node, dict) else node


RAY_TRACER_CONF = {"object_bounding_sphere": 1.0, "sdf_threshold": 5.0e-5, "line_search_step": 0.5,
                   "line_step_iters": 3, "sphere_tracing_iters": 10, "n_steps": 100, "n_secant_steps": 8}


def make_conf(embed_type, multires, log2T, base, desired, bound=1.0, view_type="NerfPos", width=512, feature=256):
    """model{} section of the reference's confs as a dict-conf (values per confs/embedder_conf_var/*)."""
    return DictConf({
        "feature_vector_size": feature,
        "implicit_network": {"d_in": 3, "d_out": 1, "dims": [width] * 8, "geometric_init": True, "bias": 0.6,
                             "skip_in": [4], "weight_norm": True, "multires": multires},
        "rendering_network": {"mode": "idr", "d_in": 9, "d_out": 3, "viewdirs_embed_type": view_type,
                              "dims": [width] * 4, "weight_norm": True, "multires_view": 4},
        "ray_tracer": dict(RAY_TRACER_CONF),
        "embedding_network": {"embed_type": embed_type, "log2_max_hash_size": log2T, "max_points_per_entry": 2,
                              "base_resolution": base, "desired_resolution": desired, "bound": bound},
    })


def quiet_build(fn, *a, **k):
    import contextlib
    import io
    import warnings
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return fn(*a, **k)
