"""ImplicitNetwork, RenderingNetwork and IDRNetwork with the reference's constructor signatures,
attribute names and state-dict keys (model/implicit_differentiable_renderer.py:11-329), so that

    train.model_class = idrk.model.implicit_differentiable_renderer.IDRNetwork

drops into the reference's IDR training loop.  Every Linear / activation pair runs in the tcgen05
contraction kernel (idrk.mlp), the encoders in the hash-grid / filter-bank kernels and ray tracing in
the ray-state kernels; torch only carries tensors, the autograd tape and O(rays) glue.
"""
import numpy as np
import torch
import torch.nn as nn

from .. import kernels as K
from .. import mlp
from ..utils import rend_util
from .custom_embedder_decoder import Custom_Embedding_Network
from .density_net import LaplaceDensity
from .embeddings.frequency_enc import SHEncoder, get_embedder
from .ray_tracing import RayTracing
from .sample_network import SampleNetwork


def _geometric_init(lin, l, n_layers, dims, out_dim, bias, multires, skip_in):
    """Sphere initialisation of the SDF MLP (IGR / reference :64-78)."""
    if l == n_layers - 2:
        torch.nn.init.normal_(lin.weight, mean=np.sqrt(np.pi) / np.sqrt(dims[l]), std=0.0001)
        torch.nn.init.constant_(lin.bias, -bias)
        return
    torch.nn.init.constant_(lin.bias, 0.0)
    std = np.sqrt(2) / np.sqrt(out_dim)
    if multires > 0 and l == 0:
        torch.nn.init.constant_(lin.weight[:, 3:], 0.0)
        torch.nn.init.normal_(lin.weight[:, :3], 0.0, std)
    elif multires > 0 and l in skip_in:
        torch.nn.init.normal_(lin.weight, 0.0, std)
        torch.nn.init.constant_(lin.weight[:, -(dims[0] - 3):], 0.0)
    else:
        torch.nn.init.normal_(lin.weight, 0.0, std)


class ImplicitNetwork(nn.Module):
    def __init__(self, feature_vector_size, d_in, d_out, dims, geometric_init=True, bias=1.0, skip_in=(),
                 weight_norm=True, multires=0, embed_type=None, log2_max_hash_size=10, max_points_per_entry=2,
                 base_resolution=64, desired_resolution=None, bound: float = 1.0):
        super().__init__()
        dims = [d_in] + list(dims) + [d_out + feature_vector_size]
        self.embed_fn = None
        self.embed_type = embed_type
        self.multires = multires
        self.dencity_net = LaplaceDensity(params_init={'beta': 0.9})
        if embed_type and multires > 0:
            self.embed_model = Custom_Embedding_Network(input_dims=d_in, network_dims=dims, embed_type=embed_type,
                                                        multires=multires, log2_max_hash_size=log2_max_hash_size,
                                                        max_points_per_entry=max_points_per_entry,
                                                        base_resolution=base_resolution,
                                                        desired_resolution=desired_resolution, bound=bound)
            self.embed_fn = self.embed_model.forward
            dims[0] = self.embed_model.embeddings_dim
        self.num_layers = len(dims)
        self.skip_in = tuple(skip_in)
        for l in range(self.num_layers - 1):
            out_dim = dims[l + 1] - dims[0] if (l + 1) in self.skip_in else dims[l + 1]
            lin = nn.Linear(dims[l], out_dim)
            if geometric_init:
                _geometric_init(lin, l, self.num_layers, dims, out_dim, bias, multires, self.skip_in)
            if weight_norm:
                lin = nn.utils.weight_norm(lin)
            setattr(self, "lin" + str(l), lin)
        self.softplus = nn.Softplus(beta=100)
        self._pipeline = mlp.SdfPipeline(self)

    # -- embedding -------------------------------------------------------------------------
    def _embed(self, x):
        return self.embed_fn(x) if self.embed_fn is not None else x

    def _fused_filter_bank(self):
        """The FourierFilterBanks module when the one-launch no-grad encoder (csrc/nffb.cu) covers its shape."""
        if self.embed_fn is None or self.embed_model.embed_type not in ("FFB", "StyleModNFFB"):
            return None
        ffb = self.embed_model.embedder_obj
        return ffb if K.nffb_fused_supported(ffb) else None

    def _embed_nograd(self, x):
        ffb = self._fused_filter_bank()
        if ffb is not None:
            return K.nffb_encode_fwd(ffb, x)[:, :ffb.embeddings_dim]
        return self._embed(x)

    # -- inference paths (no autograd) -----------------------------------------------------
    @torch.no_grad()
    def sdf(self, x: torch.Tensor) -> torch.Tensor:
        """SDF column only, [P].  What `lambda x: implicit_network(x)[:, 0]` computes in the reference
        (implicit_differentiable_renderer.py:257) without the 256 unused feature columns."""
        x = x.reshape(-1, x.shape[-1])
        if x.shape[0] == 0:
            return torch.empty(0, device=x.device)
        emb = K.operand(self._embed_nograd(x))
        return self._pipeline.run(emb, x.shape[0], want="sdf")

    def refresh_inference_weights(self, force: bool = False):
        """(Re)folds the weight-normalised layers into the inference buffers (and caches beta)."""
        self._pipeline.folded(force=force)
        self._pipeline.beta()

    def supports_device_count(self) -> bool:
        """True when the encoder can take its row count from device memory (no host sync in the tracer)."""
        return True

    @torch.no_grad()
    def sdf_compacted(self, pts: torch.Tensor, rows: int, m_count, out: torch.Tensor) -> None:
        """SDF of the first min(rows, *m_count) points of a contiguous [>=rows, 3] buffer, written to out[:rows]."""
        pipe = self._pipeline
        if self.embed_fn is None:
            emb = pipe._buf("emb", rows, 3, pts.device)
            K.split_into(pts, rows, 3, 1.0, emb, None, 4, 1, m_count)
        elif self.embed_model.embed_type == "HashGrid":
            grid = self.embed_model.embedder_obj
            if K.inference_fp16x2() and grid.n_features == 2:
                # one launch: hash encode straight into the fp16-pair operand (+ the skip connection's scaled copy)
                def encode(h, l, ld, pad, second):
                    K.hash_encode_f16pair(grid.spec(), pts, grid.tables(), grid.freq_encoding.B, rows, h, l, ld, pad,
                                          m_count, second)
                pipe.run(None, rows, want="sdf", m_count=m_count, out=out, encode=encode)
                return
            emb = pipe._buf("emb", rows, grid.embeddings_dim, pts.device)
            K.hash_encode_fwd(grid.spec(), pts, grid.tables(), grid.freq_encoding.B, out=emb, m_count=m_count, rows=rows)
        elif self._fused_filter_bank() is not None:
            ffb = self.embed_model.embedder_obj
            if K.inference_fp16x2() and K.nffb_pair_supported(ffb):
                # one launch: filter-bank encode straight into the fp16-pair operand (+ the skip connection's scaled copy)
                def encode(h, l, ld, pad, second):
                    K.nffb_encode_f16pair(ffb, pts, rows, h, l, ld, pad, m_count, second)
                pipe.run(None, rows, want="sdf", m_count=m_count, out=out, encode=encode)
                return
            emb = pipe._buf("emb", rows, ffb.embeddings_dim, pts.device)
            K.nffb_encode_fwd(ffb, pts, out=emb, m_count=m_count, rows=rows)
        else:
            # positional encoders (and filter banks wider than the fused kernel): evaluated on the whole (fixed-capacity) buffer with their module
            # kernels - rows beyond the device-side count hold stale points and are never consumed - so the call
            # sequence stays free of host syncs and can live in a CUDA graph
            emb = K.operand(self._embed(pts[:rows]))
        pipe.run(emb, rows, want="sdf", m_count=m_count, out=out)

    def _forward_inference(self, x):
        emb = K.operand(self._embed_nograd(x))
        return self._pipeline.run(emb, x.shape[0], want="full").clone()

    # -- reference API -----------------------------------------------------------------------
    def forward(self, input, compute_grad=False):
        if not torch.is_grad_enabled():
            if input.shape[0] == 0:
                return torch.empty(0, getattr(self, "lin%d" % (self.num_layers - 2)).bias.shape[0], device=input.device)
            return self._forward_inference(input)
        emb = self._embed(input)
        x = emb
        n_lin = self.num_layers - 1
        for l in range(n_lin):
            lin = getattr(self, "lin" + str(l))
            if l in self.skip_in:
                x = torch.cat([x, emb], 1) / np.sqrt(2)
            W = mlp.layer_weight(lin)
            if l < n_lin - 1:
                x = mlp.linear_act(x, W, lin.bias, "softplus", 100.0)
            else:
                x = mlp.linear(x, W, lin.bias)
        if x.is_cuda and x.dtype == torch.float32:
            # SDF squash of column 0 with the Laplace density held constant (:108-113): one kernel, differentiable twice
            return mlp.sdf_squash_rows(x, self._pipeline.beta())
        s = x[:, 0]
        rho = self.dencity_net(s.detach())
        s = torch.tanh(s / (2 + rho))
        return torch.cat([s.unsqueeze(1), x[:, 1:]], dim=1)

    def forward_with_gradient(self, x):
        """(forward(x), d sdf / d x) from ONE network evaluation.  The reference evaluates `implicit_network(p)` and
        `implicit_network.gradient(p)` separately on the same points (:321-324); `gradient` runs the same forward
        internally, so sharing it yields identical values for half the work."""
        x.requires_grad_(True)
        with torch.enable_grad():
            out = self.forward(x)
            y = out[:, :1]
            g = torch.autograd.grad(outputs=y, inputs=x, grad_outputs=torch.ones_like(y), create_graph=True,
                                    retain_graph=True, only_inputs=True)[0]
        return out, g.unsqueeze(1)

    def gradient(self, x):
        """d sdf / d x, [P, 1, 3], recorded on the tape (create_graph) like the reference (:116-128)."""
        return self.forward_with_gradient(x)[1]


class RenderingNetwork(nn.Module):
    def __init__(self, feature_vector_size, mode, d_in, d_out, dims, weight_norm=True, multires_view=0,
                 viewdirs_embed_type='NerfPos'):
        super().__init__()
        self.feature_vector_size = feature_vector_size
        self.mode = mode
        dims = [d_in + feature_vector_size] + list(dims) + [d_out]
        self.multires_view = multires_view
        self.d_in = d_in
        self.embedview_fn = None
        deep = ('HashGrid', 'FFB', 'StyleModNFFB', 'FourierFeatures', 'HashGridCUDA', 'FFBTcnn', 'HashGridTcnn')
        if viewdirs_embed_type == 'SHEncoder':
            if multires_view > 0 and mode == 'idr':
                enc = SHEncoder(3, degree=multires_view)
                self.sh_encoder = enc
                self.embedview_fn = enc.forward
                dims[0] += enc.embeddings_dim - 3
        elif viewdirs_embed_type == 'NerfPos':
            if multires_view > 0 and mode == 'idr':
                self.embedview_fn, input_ch = get_embedder(multires_view)
                dims[0] += input_ch
        elif viewdirs_embed_type in deep:
            if multires_view > 0 and mode == 'idr':
                self.embed_model = Custom_Embedding_Network(input_dims=3, network_dims=dims,
                                                            embed_type=viewdirs_embed_type, multires=multires_view,
                                                            max_points_per_entry=2,
                                                            log2_max_hash_size=multires_view - 1, base_resolution=16,
                                                            desired_resolution=512, bound=1.0)
                self.embedview_fn = self.embed_model.forward
                dims[0] += self.embed_model.embeddings_dim - 3
        else:
            raise ValueError('No Embedding Network config provided for VIEWDIRS')
        self.num_layers = len(dims)
        for l in range(self.num_layers - 1):
            lin = nn.Linear(dims[l], dims[l + 1])
            if weight_norm:
                lin = nn.utils.weight_norm(lin)
            setattr(self, "lin" + str(l), lin)
        self.relu = nn.ReLU()
        self.tanh = nn.Tanh()

    def forward(self, points, normals, view_dirs, feature_vectors):
        if self.embedview_fn is not None:
            view_dirs = self.embedview_fn(view_dirs)
        if self.mode == 'idr':
            x = torch.cat([points, view_dirs, normals, feature_vectors], dim=-1)
        elif self.mode == 'no_view_dir':
            x = torch.cat([points, normals, feature_vectors], dim=-1)
        elif self.mode == 'no_normal':
            x = torch.cat([points, view_dirs, feature_vectors], dim=-1)
        else:
            raise ValueError("unknown rendering mode %r" % self.mode)
        n_lin = self.num_layers - 1
        for l in range(n_lin):
            lin = getattr(self, "lin" + str(l))
            W = mlp.layer_weight(lin)
            x = mlp.linear_act(x, W, lin.bias, "relu" if l < n_lin - 1 else "tanh")
        return x


class IDRNetwork(nn.Module):
    def __init__(self, conf):
        super().__init__()
        self.feature_vector_size = conf.get_int('feature_vector_size')
        implicit_kwargs = dict(conf.get_config('implicit_network'))
        embedding_conf = conf.get_config('embedding_network')
        if embedding_conf is not None:
            implicit_kwargs.update(dict(embedding_conf))
        self.implicit_network = ImplicitNetwork(self.feature_vector_size, **implicit_kwargs)
        self.rendering_network = RenderingNetwork(self.feature_vector_size, **dict(conf.get_config('rendering_network')))
        self.ray_tracer = RayTracing(**dict(conf.get_config('ray_tracer')))
        self.sample_network = SampleNetwork()
        self.object_bounding_sphere = conf.get_float('ray_tracer.object_bounding_sphere')
        # host-side RNG draws (eikonal points, min-SDF steps) may be injected for exact parity runs
        self.injected_eikonal_points = None
        self.sample_generator = None            # optional torch.Generator for the host draws (per-rank in data-parallel runs)

    def forward(self, input):
        return self.shade(self.trace(input))

    def trace(self, input):
        """Ray generation + RayTracing under no_grad (reference :243-261).  Returns what the differentiable
        part needs: ray_dirs [B,N,3], cam_loc [B,3], dists [B*N], network_object_mask, object_mask."""
        intrinsics, uv, pose = input["intrinsics"], input["uv"], input["pose"]
        object_mask = input["object_mask"].reshape(-1)
        sphere = None
        if rend_util._kernel_path(uv, pose, intrinsics):
            # rays and their bounding-sphere intersections from one launch (no boolean indexing, no host sync)
            ray_dirs, cam_loc, t_sph, hit = rend_util.camera_rays_and_sphere(uv, pose, intrinsics, self.object_bounding_sphere)
            sphere = (t_sph, hit)
        else:
            ray_dirs, cam_loc = rend_util.get_camera_params(uv, pose, intrinsics)
        self.implicit_network.eval()
        with torch.no_grad():
            self.ray_tracer.train(self.training)
            _, network_object_mask, dists = self.ray_tracer(sdf=self.implicit_network.sdf, cam_loc=cam_loc,
                                                            object_mask=object_mask, ray_directions=ray_dirs,
                                                            sphere_intersections=sphere)
        self.implicit_network.train()
        return {"ray_dirs": ray_dirs, "cam_loc": cam_loc, "dists": dists, "network_object_mask": network_object_mask,
                "object_mask": object_mask}

    def shade(self, traced, eikonal_points=None):
        """Differentiable part of the forward (reference :262-319) given the traced distances."""
        with mlp.shared_weights():
            return self._shade(traced, eikonal_points)

    def _shade(self, traced, eikonal_points=None):
        ray_dirs, cam_loc, dists = traced["ray_dirs"], traced["cam_loc"], traced["dists"]
        network_object_mask, object_mask = traced["network_object_mask"], traced["object_mask"]
        batch_size, num_pixels, _ = ray_dirs.shape
        device = ray_dirs.device
        points = (cam_loc.unsqueeze(1) + dists.reshape(batch_size, num_pixels, 1) * ray_dirs).reshape(-1, 3)
        ray_dirs = ray_dirs.reshape(-1, 3)

        if self.training:
            if eikonal_points is None:
                eikonal_points = self._draw_eikonal(batch_size * num_pixels, device)
            cam_rep = cam_loc.unsqueeze(1).repeat(1, num_pixels, 1).reshape(-1, 3)
            rgb_values, grad_theta, sdf_output = self.render_training(points, dists, ray_dirs, cam_rep,
                                                                      network_object_mask & object_mask, eikonal_points)
        else:
            sdf_output = self.implicit_network(points)[:, 0:1]
            surface_mask = network_object_mask
            differentiable_surface_points = points[surface_mask]
            grad_theta = None
            view = -ray_dirs[surface_mask]
            rgb_values = torch.ones_like(points).float()
            if differentiable_surface_points.shape[0] > 0:
                rgb_values[surface_mask] = self.get_rbg_value(differentiable_surface_points, view)

        return {'points': points, 'rgb_values': rgb_values, 'sdf_output': sdf_output,
                'network_object_mask': network_object_mask, 'object_mask': object_mask, 'grad_theta': grad_theta}

    def _draw_eikonal(self, n_rays, device):
        if self.injected_eikonal_points is not None:
            return self.injected_eikonal_points.to(device)
        r = self.object_bounding_sphere      # drawn on the host generator, like the reference (:279)
        return torch.empty(n_rays // 2, 3).uniform_(-r, r, generator=self.sample_generator).to(device)

    def render_training(self, points, dists, ray_dirs, cam_rep, surface_mask, eik):
        """Training branch of the reference forward (:264-308) on FIXED shapes, without its duplicate evaluations.

        The reference gathers the N_s surface rays with boolean masks and evaluates the network on `points`,
        on `surface_points`, and on [surface | eikonal | points] for the gradient.  Surface points are a subset of
        `points`, and `gradient()` runs a forward pass internally, so the same numbers come out of ONE
        forward-with-gradient on [eikonal | points]: its SDF column on the `points` rows is `sdf_output`, its gradient
        rows are `grad_theta`, and the surface rows of both are read out by mask.  Sample network + rendering then
        run on all N rays with the non-surface rows masked afterwards (their rgb is the constant 1 and they receive
        zero gradient, exactly as in the reference).  Fixed shapes keep the step free of host syncs and capturable in
        a CUDA graph.  When `points` itself carries gradient (trainable cameras) the separate evaluation of the
        reference is kept so that d sdf / d pose is preserved."""
        n_eik = eik.shape[0]
        if points.requires_grad:
            sdf_output = self.implicit_network(points)[:, 0:1]
            g = self.implicit_network.gradient(torch.cat([eik, points.clone().detach()], 0))
        else:
            out_all, g = self.implicit_network.forward_with_gradient(torch.cat([eik, points], 0))
            sdf_output = out_all[n_eik:, 0:1]
        grad_theta = g[:, 0, :]
        normals0 = g[n_eik:, 0, :].clone().detach()
        sdf0 = sdf_output.detach()
        mask = surface_mask.unsqueeze(-1)
        denom = (normals0 * ray_dirs.detach()).sum(dim=-1, keepdim=True)
        denom = torch.where(mask, denom, torch.ones_like(denom))
        t_theta = dists.unsqueeze(-1) - (sdf_output - sdf0) / denom           # sample_network.py:10-20
        diff_points = cam_rep + t_theta * ray_dirs
        rgb = self.get_rbg_value(diff_points, -ray_dirs)
        rgb_values = torch.where(mask, rgb, torch.ones_like(rgb))
        return rgb_values, grad_theta, sdf_output

    def get_rbg_value(self, points, view_dirs):
        output, g = self.implicit_network.forward_with_gradient(points)
        normals = g[:, 0, :]
        feature_vectors = output[:, 1:]
        return self.rendering_network(points, normals, view_dirs, feature_vectors)
