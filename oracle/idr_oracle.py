"""CPU oracle for the IDR hash-grid rendering hot path  --  TEST INFRASTRUCTURE ONLY.

This file restates, in plain torch-CPU / numpy arithmetic, what the reference
(ArtoriasAbyssslayer/HashModNFFBanks-IDR, tree at /root/reference/code) computes on the
path named by BASELINE.json:north_star.  It is the *checker*: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.
The product (hashmodnffbanks-idr_b200/) never does; it fails loudly without its CUDA library.

Parity pinning: the reference ships no golden vectors, so this oracle is pinned against
outputs of the reference itself, generated in the build container by
tests/golden/make_golden.py (imports /root/reference/code) and committed under
tests/golden/*.npz.  tests/test_oracle_golden.py checks every function below against them.

All functions work on a flat ``sd`` dict that uses the reference's state_dict key names
 so weights move freely between reference, oracle and product.
Reference line numbers are cited per function as  file:line.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor

PRIME_Y = 3                # model/embeddings/hashGridEmbedding.py:14  (HASH_PRIMES[1])
PRIME_Z = 2654435761       # model/embeddings/hashGridEmbedding.py:14  (HASH_PRIMES[2])


# --------------------------------------------------------------------------------------
# Multi-resolution hash grid                                  hashGridEmbedding.py:32-155
# --------------------------------------------------------------------------------------
def level_schedule(n_levels: int, base_res: int, desired_res: int, log2_T: int,
                   in_dim: int = 3) -> Tuple[List[int], List[int]]:
    """Per-level grid resolution and table rows.  hashGridEmbedding.py:125-138
    (Python doubles, floor(base * growth**l), T = min(res**d, 2**log2_T))."""
    growth = math.exp((math.log(desired_res) - math.log(base_res)) / (n_levels - 1))
    res, rows = [], []
    for lvl in range(n_levels):
        r = math.floor(base_res * (growth ** lvl))
        res.append(r)
        rows.append(min(r ** in_dim, 2 ** log2_T))
    return res, rows


def fourier_sigma(base_res: int, desired_res: int) -> float:
    """sigma of the Fourier prefix at construction.  hashGridEmbedding.py:141"""
    return (math.log(desired_res) - math.log(base_res)) / (base_res - 1)


def corner_hash(x: Tensor, res: int, rows: int, floor_mode: bool = False) -> np.ndarray:
    """uint32 hash row of the 8 cell corners of every point -> int64 array [P, 8].

    hashGridEmbedding.py:84-85 (scale by resolution, `.long()` truncates toward zero),
    :93 (corner k takes xi+1 in dimension d iff bit d of k is set, via bin_mask :76-80),
    :33-40 (multiply by primes [1, 3, 2654435761], keep 32 bits, xor-fold, mod table rows).
    `floor_mode` selects floor() instead of trunc for the trilinear extension.
    """
    xs = (x.detach().to(torch.float32) * float(res))
    xi = torch.floor(xs).to(torch.int64) if floor_mode else xs.to(torch.int64)
    xi = xi.numpy()
    out = np.empty((xi.shape[0], 8), dtype=np.int64)
    for k in range(8):
        c0 = (xi[:, 0] + ((k >> 0) & 1)).astype(np.uint64) & np.uint64(0xFFFFFFFF)
        c1 = (xi[:, 1] + ((k >> 1) & 1)).astype(np.uint64) & np.uint64(0xFFFFFFFF)
        c2 = (xi[:, 2] + ((k >> 2) & 1)).astype(np.uint64) & np.uint64(0xFFFFFFFF)
        h = (c0 ^ ((c1 * np.uint64(PRIME_Y)) & np.uint64(0xFFFFFFFF))
             ^ ((c2 * np.uint64(PRIME_Z)) & np.uint64(0xFFFFFFFF)))
        out[:, k] = (h % np.uint64(rows)).astype(np.int64)
    return out


def hash_level(x: Tensor, table: Tensor, res: int, mode: str = "reference") -> Tensor:
    """One level's features [P, F].  hashGridEmbedding.py:81-102.

    mode="reference": the reference's fractional part is `x - x.float()` == 0 (:86), so the
    eight interpolation weights are (1,0,...,0) and the result is exactly the floor-corner row.
    mode="trilinear": the intended interpolation (floor / frac, 8 weighted corners) - our
    documented extension, not a parity target.
    """
    rows = table.shape[0]
    if mode == "reference":
        idx = torch.from_numpy(corner_hash(x, res, rows)[:, 0])
        return table[idx]
    idx = torch.from_numpy(corner_hash(x, res, rows, floor_mode=True))      # [P, 8]
    xs = x.to(torch.float32) * float(res)
    fr = xs - torch.floor(xs)                                               # [P, 3]
    feats = torch.zeros(x.shape[0], table.shape[1], dtype=torch.float32)
    for k in range(8):
        w = torch.ones(x.shape[0], dtype=torch.float32)
        for d in range(3):
            w = w * (fr[:, d] if (k >> d) & 1 else (1.0 - fr[:, d]))
        feats = feats + w[:, None] * table[idx[:, k]]
    return feats


# --------------------------------------------------------------------------------------
# tiny-cuda-nn grid semantics (HashGridTcnn / FFBTcnn)      tcnn_src/hashGridEncoderTcnn.py:63-80
# --------------------------------------------------------------------------------------
# PARITY UNPINNED for this block: the arithmetic lives in tiny-cuda-nn (NVlabs), which the reference installs from
# git HEAD (README.md:58-60; absent from requirements.txt / environment.yml) and which is not present under
# /root/reference nor importable here; the reference holds no test or vector at that boundary.  What follows restates
# tiny-cuda-nn's published Grid encoding (include/tiny-cuda-nn/encodings/grid.h: grid_scale, grid_resolution,
# pos_fract, grid_index, coherent prime hash; GridEncodingTemplated ctor for the per-level parameter counts) for
# otype "Grid", type "Hash", interpolation "Linear", as configured at tcnn_src/hashGridEncoderTcnn.py:63-80.
NGP_PRIME_Y = 2654435761
NGP_PRIME_Z = 805459861


def ngp_level_layout(n_levels: int, log2_T: int, base_res: int, per_level_scale: float):
    """(scale_l, R_l, rows_l, row offsets) per level: scale = exp2f(l * log2f(s)) * base - 1, R = ceil(scale) + 1,
    rows = min(next_multiple(R^3, 8), 2^log2_T)."""
    log2s = np.float32(math.log2(per_level_scale))
    scales, res, rows, offs = [], [], [], [0]
    for l in range(n_levels):
        sc = np.float32(np.exp2(np.float32(l) * log2s) * np.float32(base_res) - np.float32(1.0))
        R = int(np.ceil(sc)) + 1
        n = min(R ** 3, (2 ** 32 - 1) // 2)
        n = min(((n + 7) // 8) * 8, 1 << log2_T)
        scales.append(float(sc)); res.append(R); rows.append(n); offs.append(offs[-1] + n)
    return scales, res, rows, offs


def ngp_corner_rows(x: Tensor, scale: float, R: int, rows: int) -> Tuple[np.ndarray, Tensor]:
    """(rows [P, 8] int64, frac [P, 3]) of one level: pos = fma(x, scale, 0.5); vertex = floor(pos) (+1 per set corner
    bit, bit d of k <-> dimension d); dense stride index while R^3 <= rows, else the coherent prime hash; mod rows."""
    x32 = x.detach().to(torch.float32).numpy()
    pos = (x32.astype(np.float64) * np.float64(np.float32(scale)) + 0.5).astype(np.float32)     # fma: one rounding
    fl = np.floor(pos)
    frac = torch.from_numpy(pos - fl)
    g = fl.astype(np.int64)
    dense = R ** 3 <= rows
    out = np.empty((x32.shape[0], 8), dtype=np.int64)
    M = np.uint64(0xFFFFFFFF)
    for k in range(8):
        c = [(g[:, d] + ((k >> d) & 1)).astype(np.uint64) & M for d in range(3)]
        if dense:
            idx = (c[0] + ((c[1] * np.uint64(R)) & M) + ((c[2] * np.uint64(R * R)) & M)) & M
        else:
            idx = c[0] ^ ((c[1] * np.uint64(NGP_PRIME_Y)) & M) ^ ((c[2] * np.uint64(NGP_PRIME_Z)) & M)
        out[:, k] = (idx % np.uint64(rows)).astype(np.int64)
    return out, frac


def ngp_grid_encode(x: Tensor, params: Tensor, n_levels: int, n_feat: int, log2_T: int, base_res: int,
                    per_level_scale: float = 2.0) -> Tensor:
    """tcnn.Encoding("Grid", "Hash", "Linear")(x) for x in [0, 1]^3 -> [P, L * F]; differentiable in `params` (the flat
    parameter vector, level-major, rows of F) - the gradient w.r.t. x follows from the trilinear weights
    (d/dx = scale * sum_k dw_k/dfrac * row_k) and is returned by ngp_grid_dx."""
    scales, res, rows, offs = ngp_level_layout(n_levels, log2_T, base_res, per_level_scale)
    feats = []
    for l in range(n_levels):
        table = params[offs[l] * n_feat: offs[l + 1] * n_feat].view(rows[l], n_feat)
        idx, fr = ngp_corner_rows(x, scales[l], res[l], rows[l])
        idx = torch.from_numpy(idx)
        acc = torch.zeros(x.shape[0], n_feat, dtype=torch.float32)
        for k in range(8):
            w = torch.ones(x.shape[0], dtype=torch.float32)
            for d in range(3):
                w = w * (fr[:, d] if (k >> d) & 1 else (1.0 - fr[:, d]))
            acc = acc + w[:, None] * table[idx[:, k]]
        feats.append(acc)
    return torch.cat(feats, dim=-1)


def ngp_grid_dx(x: Tensor, params: Tensor, dy: Tensor, n_levels: int, n_feat: int, log2_T: int, base_res: int,
                per_level_scale: float = 2.0) -> Tensor:
    """dL/dx [P, 3] of ngp_grid_encode for upstream dy [P, L * F] (piecewise: valid inside a cell)."""
    scales, res, rows, offs = ngp_level_layout(n_levels, log2_T, base_res, per_level_scale)
    dx = torch.zeros(x.shape[0], 3, dtype=torch.float64)
    for l in range(n_levels):
        table = params.detach()[offs[l] * n_feat: offs[l + 1] * n_feat].view(rows[l], n_feat).double()
        idx, fr = ngp_corner_rows(x, scales[l], res[l], rows[l])
        idx, fr = torch.from_numpy(idx), fr.double()
        g = dy[:, l * n_feat:(l + 1) * n_feat].double()
        for k in range(8):
            dot = (table[idx[:, k]] * g).sum(-1)
            for d in range(3):
                w = torch.ones(x.shape[0], dtype=torch.float64)
                for e in range(3):
                    if e == d:
                        w = w * (1.0 if (k >> e) & 1 else -1.0)
                    else:
                        w = w * (fr[:, e] if (k >> e) & 1 else (1.0 - fr[:, e]))
                dx[:, d] += w * dot * float(np.float32(scales[l]))
    return dx.float()


def fourier_feature(x: Tensor, B: Tensor, include_input: bool = True) -> Tensor:
    """[x | sin(2*pi*x@B) | cos(2*pi*x@B)].  frequency_enc.py:63-67"""
    proj = torch.matmul(2 * np.pi * x, B)
    parts = [torch.sin(proj), torch.cos(proj)]
    return torch.cat(([x] if include_input else []) + parts, dim=-1)


def hashgrid_tables(sd: Dict[str, Tensor], prefix: str, n_levels: int) -> List[Tensor]:
    return [sd[f"{prefix}levels.{l}.embedding.weight"] for l in range(n_levels)]


def hashgrid_embed(x: Tensor, sd: Dict[str, Tensor], prefix: str, n_levels: int, base_res: int,
                   desired_res: int, mode: str = "reference") -> Tensor:
    """MultiResHashGridMLP.forward with include_input=True.  hashGridEmbedding.py:150-153:
    cat([FourierFeature(x), level_0(x), ..., level_{L-1}(x)])  ->  [P, 3 + 2L + L*F]."""
    tables = hashgrid_tables(sd, prefix, n_levels)
    log2_T = 62  # table row counts come from the tensors themselves
    res, _ = level_schedule(n_levels, base_res, desired_res, log2_T)
    pre = fourier_feature(x, sd[f"{prefix}freq_encoding.B"])
    lvls = [hash_level(x, tables[l], res[l], mode) for l in range(n_levels)]
    return torch.cat([pre] + lvls, dim=-1)


# --------------------------------------------------------------------------------------
# Positional encoding                                              frequency_enc.py:6-51
# --------------------------------------------------------------------------------------
def positional_encoding(x: Tensor, num_freqs: int, max_freq_log2: float,
                        include_input: bool = True) -> Tensor:
    """PositionalEncoding.embed.  frequency_enc.py:19-51.  Note the quirk: with include_input
    the input appears TWICE (once from embed_fns :24, once more in embed :46-47)."""
    bands = 2.0 ** torch.linspace(0.0, max_freq_log2, num_freqs)
    cols = [x] if include_input else []
    for f in bands:
        cols.append(torch.sin(x * f))
        cols.append(torch.cos(x * f))
    enc = torch.cat(cols, dim=-1)
    return torch.cat([x, enc], dim=-1) if include_input else enc


def view_embed_nerfpos(v: Tensor, multires_view: int) -> Tensor:
    """get_embedder(multires).  frequency_enc.py:156-168 -> PositionalEncoding.embed with
    max_freq_log2 = multires-1 (emits d*(2+2*Nf) columns)."""
    return positional_encoding(v, multires_view, multires_view - 1, True)


# --------------------------------------------------------------------------------------
# NFFB / StyleModNFFB                                                  nffb3d.py:26-194
# --------------------------------------------------------------------------------------
def style_attention(style: Tensor, sd: Dict[str, Tensor], prefix: str) -> Tensor:
    """StyleAttention.forward.  styleMod.py:26-44.  softmax over a size-1 dim is 1.0 and
    InstanceNorm1d on a 2-D tensor normalises each row (biased variance, eps 1e-5)."""
    lin = F.linear(style, sd[f"{prefix}linear_transform.weight"], sd[f"{prefix}linear_transform.bias"])
    mu = lin.mean(dim=1, keepdim=True)
    var = lin.var(dim=1, unbiased=False, keepdim=True)
    return (lin - mu) / torch.sqrt(var + 1e-5)


def nffb_embed(p: Tensor, sd: Dict[str, Tensor], prefix: str, n_levels: int, n_feat: int,
               base_res: int, desired_res: int, bound: float, style: bool,
               mode: str = "reference") -> Tensor:
    """FourierFilterBanks.forward (has_out=False, SIREN, PositionalEncodingNET).
    nffb3d.py:122-194;  w0 = L**F - L (:83);  output [u | sum(features)/L]  (:190-193)."""
    L = n_levels
    z = p / bound                                                       # :131
    u = (p + bound) / (2 * bound)                                       # :132
    grid = hashgrid_embed(u, sd, f"{prefix}grid_enc.", L, base_res, desired_res, mode)
    g = grid[:, 3:].reshape(-1, L, 2 * n_feat)                          # :137-138
    enc = [positional_encoding(g[:, i, :], L, L - 1, True) for i in range(L)]   # :142-144
    w0 = float(L ** n_feat - L)
    total = None
    for j in range(L - 1):                                              # :163
        z = torch.sin(w0 * F.linear(z, sd[f"{prefix}ff_lin{j}.weight"], sd[f"{prefix}ff_lin{j}.bias"]))
        if j > 0:
            e = enc[j - 1]
            if style:
                e = style_attention(e, sd, f"{prefix}StyleAttentionBlock.")
            e = e + z
            o = F.linear(e, sd[f"{prefix}out_layer.weight"], sd[f"{prefix}out_layer.bias"])
            total = o if total is None else total + o
    return torch.cat([u, total / L], dim=-1)


# --------------------------------------------------------------------------------------
# Embedder selector                                      custom_embedder_decoder.py:13-164
# --------------------------------------------------------------------------------------
class EmbedCfg:
    """The subset of Custom_Embedding_Network's arguments that changes the arithmetic."""

    def __init__(self, embed_type: str, multires: int, log2_max_hash_size: int,
                 max_points_per_entry: int, base_resolution: int, desired_resolution: int,
                 bound: float = 1.0, mode: str = "reference"):
        self.embed_type = embed_type
        self.multires = multires
        self.log2_T = log2_max_hash_size
        self.n_feat = max_points_per_entry
        self.base = base_resolution
        self.desired = desired_resolution
        self.bound = bound
        self.mode = mode

    def width(self) -> int:
        L, Fd = self.multires, self.n_feat
        if self.embed_type == "HashGrid":
            return 3 + 2 * L + L * Fd
        if self.embed_type in ("FFB", "StyleModNFFB"):
            return 3 + 2 * (Fd * (2 + 2 * L))
        if self.embed_type == "NerfPos":
            return 3 * (2 + 2 * L)
        if self.embed_type == "FourierFeatures":
            return 9
        raise ValueError("Not a valid embedding model type")


def embed(x: Tensor, sd: Dict[str, Tensor], prefix: str, cfg: EmbedCfg) -> Tensor:
    """Custom_Embedding_Network.forward.  custom_embedder_decoder.py:147-164."""
    pre = f"{prefix}embedder_obj."
    if cfg.embed_type == "HashGrid":
        return hashgrid_embed(x, sd, pre, cfg.multires, cfg.base, cfg.desired, cfg.mode)
    if cfg.embed_type in ("FFB", "StyleModNFFB"):
        return nffb_embed(x, sd, pre, cfg.multires, cfg.n_feat, cfg.base, cfg.desired, cfg.bound,
                          cfg.embed_type == "StyleModNFFB", cfg.mode)
    if cfg.embed_type == "NerfPos":             # max_freq_log2 = log2_max_hash_size (:74-81)
        return positional_encoding(x, cfg.multires, cfg.log2_T, True)
    if cfg.embed_type == "FourierFeatures":
        return fourier_feature(x, sd[f"{pre}B"])
    raise ValueError("Not a valid embedding model type")


# --------------------------------------------------------------------------------------
# MLPs                                       implicit_differentiable_renderer.py:11-223
# --------------------------------------------------------------------------------------
def effective_weight(sd: Dict[str, Tensor], name: str) -> Tensor:
    """Legacy weight_norm: W = g * v / ||v||_row  (nn.utils.weight_norm, dim=0), or the plain
    weight when the layer is not weight-normalised.  implicit_differentiable_renderer.py:80-81."""
    if f"{name}.weight_g" in sd:
        return torch._weight_norm(sd[f"{name}.weight_v"], sd[f"{name}.weight_g"], 0)
    return sd[f"{name}.weight"]


def laplace_density(s: Tensor, beta_param: Tensor) -> Tensor:
    """LaplaceDensity.density_func evaluated without gradient.  density_net.py:20-30."""
    with torch.no_grad():
        beta = beta_param.abs() + 1e-4
        return (1.0 / beta) * (0.5 + 0.5 * s.sign() * torch.expm1(-s.abs() / beta))


def implicit_forward(x: Tensor, sd: Dict[str, Tensor], cfg: Optional[EmbedCfg], n_lin: int = 9,
                     skip_in: Sequence[int] = (4,), prefix: str = "implicit_network.") -> Tensor:
    """ImplicitNetwork.forward -> [P, 1 + feature].  implicit_differentiable_renderer.py:89-113:
    embed, n_lin weight-normalised Linear layers with Softplus(beta=100) between them, the skip
    concat /sqrt(2) before layer 4, and the in-place squash of column 0 (:112)."""
    emb = embed(x, sd, f"{prefix}embed_model.", cfg) if cfg is not None else x
    h = emb
    for l in range(n_lin):
        if l in skip_in:
            h = torch.cat([h, emb], dim=1) / np.sqrt(2)
        h = F.linear(h, effective_weight(sd, f"{prefix}lin{l}"), sd[f"{prefix}lin{l}.bias"])
        if l < n_lin - 1:
            h = F.softplus(h, beta=100)
    s = h[:, 0]
    rho = laplace_density(s.detach(), sd[f"{prefix}dencity_net.beta"])
    s = torch.tanh(s / (2 + rho))
    return torch.cat([s[:, None], h[:, 1:]], dim=1)


def implicit_gradient(x: Tensor, sd, cfg, **kw) -> Tensor:
    """ImplicitNetwork.gradient -> [P, 1, 3] with create_graph.  :116-128"""
    x.requires_grad_(True)
    y = implicit_forward(x, sd, cfg, **kw)[:, :1]
    g = torch.autograd.grad(y, x, torch.ones_like(y), create_graph=True, retain_graph=True)[0]
    return g.unsqueeze(1)


def rendering_forward(points: Tensor, normals: Tensor, view_dirs: Tensor, feats: Tensor,
                      sd: Dict[str, Tensor], multires_view: int = 4, n_lin: int = 5,
                      view_cfg: Optional[EmbedCfg] = None,
                      prefix: str = "rendering_network.") -> Tensor:
    """RenderingNetwork.forward (mode 'idr').  implicit_differentiable_renderer.py:202-223."""
    if view_cfg is not None:
        v = embed(view_dirs, sd, f"{prefix}embed_model.", view_cfg)
    elif multires_view > 0:
        v = view_embed_nerfpos(view_dirs, multires_view)
    else:
        v = view_dirs
    h = torch.cat([points, v, normals, feats], dim=-1)
    for l in range(n_lin):
        h = F.linear(h, effective_weight(sd, f"{prefix}lin{l}"), sd[f"{prefix}lin{l}.bias"])
        if l < n_lin - 1:
            h = torch.relu(h)
    return torch.tanh(h)


# --------------------------------------------------------------------------------------
# Camera / sphere                                                rend_util.py:48-162
# --------------------------------------------------------------------------------------
def camera_rays(uv: Tensor, pose: Tensor, K: Tensor) -> Tuple[Tensor, Tensor]:
    """get_camera_params + lift for a 4x4 pose.  rend_util.py:48-75, :87-100."""
    cam_loc = pose[:, :3, 3]
    fx, fy = K[:, 0, 0, None], K[:, 1, 1, None]
    cx, cy, sk = K[:, 0, 2, None], K[:, 1, 2, None], K[:, 0, 1, None]
    xs, ys = uv[:, :, 0], uv[:, :, 1]
    z = torch.ones_like(xs)
    xl = (xs - cx + cy * sk / fy - sk * ys / fy) / fx * z
    yl = (ys - cy) / fy * z
    pix = torch.stack((xl, yl, z, torch.ones_like(z)), dim=-1).permute(0, 2, 1)
    world = torch.bmm(pose, pix).permute(0, 2, 1)[:, :, :3]
    dirs = F.normalize(world - cam_loc[:, None, :], dim=2)
    return dirs, cam_loc


def sphere_intersection(cam_loc: Tensor, dirs: Tensor, r: float = 1.0) -> Tuple[Tensor, Tensor]:
    """get_sphere_intersection.  rend_util.py:141-162."""
    n_img, n_pix, _ = dirs.shape
    dot = torch.bmm(dirs, cam_loc.unsqueeze(-1)).squeeze(-1)
    under = (dot ** 2 - (cam_loc.norm(2, 1, keepdim=True) ** 2 - r ** 2)).reshape(-1)
    hit = under > 0
    t = torch.zeros(n_img * n_pix, 2)
    t[hit] = torch.sqrt(under[hit]).unsqueeze(-1) * torch.tensor([-1.0, 1.0])
    t[hit] -= dot.reshape(-1)[hit].unsqueeze(-1)
    t = t.reshape(n_img, n_pix, 2).clamp_min(0.0)
    return t, hit.reshape(n_img, n_pix)


# --------------------------------------------------------------------------------------
# Ray tracing                                                     ray_tracing.py:5-298
# --------------------------------------------------------------------------------------
class RayTracerOracle:
    """Literal restatement of RayTracing (all under no_grad).  ray_tracing.py:26-298.
    `min_sdf_steps` lets the caller inject the U(0,1)^n_steps vector that the reference draws
    from the CPU generator at :277 so both sides see the same randomness."""

    def __init__(self, object_bounding_sphere=1.0, sdf_threshold=5.0e-5, line_search_step=0.5,
                 line_step_iters=1, sphere_tracing_iters=10, n_steps=100, n_secant_steps=8):
        self.r = object_bounding_sphere
        self.thr = sdf_threshold
        self.ls_step = line_search_step
        self.ls_iters = line_step_iters
        self.st_iters = sphere_tracing_iters
        self.n_steps = n_steps
        self.n_secant = n_secant_steps
        self.training = True
        self.stats: Dict[str, int] = {}
        # Test bookkeeping (does not touch the results): for every ray the smallest distance by which any SDF value
        # that drove one of its discrete decisions (<= threshold :134-142, < 0 :164-165 / :212-218, t0 < t1 :185-186)
        # missed the decision boundary.  A ray whose hit/miss mask differs between two evaluators of the same SDF is
        # only legitimate when this margin is inside the evaluators' SDF tolerance.
        self.margin: Optional[Tensor] = None

    @torch.no_grad()
    def __call__(self, sdf: Callable[[Tensor], Tensor], cam_loc: Tensor, object_mask: Tensor,
                 ray_directions: Tensor, min_sdf_steps: Optional[Tensor] = None):
        B, N, _ = ray_directions.shape
        self.stats = {"sdf_calls": 0, "sdf_points": 0}

        def f(pts):
            self.stats["sdf_calls"] += 1
            self.stats["sdf_points"] += int(pts.shape[0])
            return sdf(pts)

        self.margin = torch.full((B * N,), float("inf"))
        t_sph, hit = sphere_intersection(cam_loc, ray_directions, self.r)
        pts, unf_start, t0, t1, min_dis, max_dis = self._sphere_trace(B, N, f, cam_loc, ray_directions, hit, t_sph)
        net_mask = t0 < t1                                                   # :39
        samp_mask = unf_start
        samp_net = torch.zeros_like(samp_mask)
        if samp_mask.sum() > 0:                                              # :44-59
            mm = torch.zeros(B, N, 2)
            mm.reshape(-1, 2)[samp_mask, 0] = t0[samp_mask]
            mm.reshape(-1, 2)[samp_mask, 1] = t1[samp_mask]
            s_pts, samp_net, s_d = self._sampler(f, cam_loc, object_mask, ray_directions, mm, samp_mask)
            pts[samp_mask] = s_pts[samp_mask]
            t0[samp_mask] = s_d[samp_mask]
            net_mask[samp_mask] = samp_net[samp_mask]
        self.stats["n_sampler"] = int(samp_mask.sum())
        if not self.training:                                                # :66-69
            return pts, net_mask, t0

        dirs = ray_directions.reshape(-1, 3)
        hit_f = hit.reshape(-1)
        in_mask = ~net_mask & object_mask & ~samp_mask                       # :74-75
        out_mask = ~object_mask & ~samp_mask
        left = (in_mask | out_mask) & ~hit_f                                 # :77-82
        cam_rep = cam_loc.unsqueeze(1).repeat(1, N, 1).reshape(-1, 3)
        if left.sum() > 0:
            t0[left] = -(dirs[left] * cam_rep[left]).sum(-1)
            pts[left] = cam_rep[left] + t0[left].unsqueeze(1) * dirs[left]
        m = (in_mask | out_mask) & hit_f                                     # :84-92
        if m.sum() > 0:
            sel = net_mask & out_mask
            min_dis[sel] = t0[sel]
            mp, md = self._min_sdf(N, f, cam_loc, dirs, m, min_dis, max_dis, min_sdf_steps)
            pts[m] = mp
            t0[m] = md
        return pts, net_mask, t0

    def _sphere_trace(self, B, N, f, cam_loc, dirs, hit, t_sph):             # :98-187
        sph_pts = cam_loc.reshape(B, 1, 1, 3) + t_sph.unsqueeze(-1) * dirs.unsqueeze(2)
        unf_s = hit.reshape(-1).clone()
        unf_e = hit.reshape(-1).clone()
        ps = torch.zeros(B * N, 3)
        ps[unf_s] = sph_pts[:, :, 0, :].reshape(-1, 3)[unf_s]
        t0 = torch.zeros(B * N)
        t0[unf_s] = t_sph.reshape(-1, 2)[unf_s, 0]
        pe = torch.zeros(B * N, 3)
        pe[unf_e] = sph_pts[:, :, 1, :].reshape(-1, 3)[unf_e]
        t1 = torch.zeros(B * N)
        t1[unf_e] = t_sph.reshape(-1, 2)[unf_e, 1]
        min_dis, max_dis = t0.clone(), t1.clone()
        nxt_s = torch.zeros_like(t0)
        nxt_s[unf_s] = f(ps[unf_s])
        nxt_e = torch.zeros_like(t1)
        nxt_e[unf_e] = f(pe[unf_e])
        self._note(unf_s, nxt_s)
        self._note(unf_e, nxt_e)
        it = 0
        while True:
            cur_s = torch.zeros_like(t0)
            cur_s[unf_s] = nxt_s[unf_s]
            cur_s[cur_s <= self.thr] = 0
            cur_e = torch.zeros_like(t1)
            cur_e[unf_e] = nxt_e[unf_e]
            cur_e[cur_e <= self.thr] = 0
            unf_s = unf_s & (cur_s > self.thr)
            unf_e = unf_e & (cur_e > self.thr)
            if (unf_s.sum() == 0 and unf_e.sum() == 0) or it == self.st_iters:
                break
            it += 1
            t0 = t0 + cur_s
            t1 = t1 - cur_e
            ps = (cam_loc.unsqueeze(1) + t0.reshape(B, N, 1) * dirs).reshape(-1, 3)
            pe = (cam_loc.unsqueeze(1) + t1.reshape(B, N, 1) * dirs).reshape(-1, 3)
            nxt_s = torch.zeros_like(t0)
            nxt_s[unf_s] = f(ps[unf_s])
            nxt_e = torch.zeros_like(t1)
            nxt_e[unf_e] = f(pe[unf_e])
            self._note(unf_s, nxt_s)
            self._note(unf_e, nxt_e)
            bad_s, bad_e = nxt_s < 0, nxt_e < 0
            k = 0
            while (bad_s.sum() > 0 or bad_e.sum() > 0) and k < self.ls_iters:
                back = (1 - self.ls_step) / (2 ** k)
                t0[bad_s] -= back * cur_s[bad_s]
                ps[bad_s] = (cam_loc.unsqueeze(1) + t0.reshape(B, N, 1) * dirs).reshape(-1, 3)[bad_s]
                t1[bad_e] += back * cur_e[bad_e]
                pe[bad_e] = (cam_loc.unsqueeze(1) + t1.reshape(B, N, 1) * dirs).reshape(-1, 3)[bad_e]
                nxt_s[bad_s] = f(ps[bad_s])
                nxt_e[bad_e] = f(pe[bad_e])
                self._note(bad_s, nxt_s)
                self._note(bad_e, nxt_e)
                bad_s, bad_e = nxt_s < 0, nxt_e < 0
                k += 1
            live = unf_s | unf_e
            self.margin[live] = torch.minimum(self.margin[live], (t1 - t0).abs()[live])
            unf_s = unf_s & (t0 < t1)
            unf_e = unf_e & (t0 < t1)
        return ps, unf_s, t0, t1, min_dis, max_dis

    def _note(self, sel: Tensor, vals: Tensor) -> None:
        """margin bookkeeping for SDF values that are compared with the threshold and with 0."""
        if self.margin is None or int(sel.sum()) == 0:
            return
        v = vals[sel]
        self.margin[sel] = torch.minimum(self.margin[sel], torch.minimum((v - self.thr).abs(), v.abs()))

    def _sampler(self, f, cam_loc, object_mask, dirs, mm, samp_mask):        # :189-249
        B, N, _ = dirs.shape
        n_tot = B * N
        out_pts = torch.zeros(n_tot, 3)
        out_d = torch.zeros(n_tot)
        lin = torch.linspace(0, 1, steps=self.n_steps).view(1, 1, -1)
        z = mm[:, :, 0].unsqueeze(-1) + lin * (mm[:, :, 1] - mm[:, :, 0]).unsqueeze(-1)
        P = cam_loc.reshape(B, 1, 1, 3) + z.unsqueeze(-1) * dirs.unsqueeze(2)
        ridx = torch.nonzero(samp_mask).flatten()
        P = P.reshape(-1, self.n_steps, 3)[samp_mask]
        z = z.reshape(-1, self.n_steps)[samp_mask]
        vals = torch.cat([f(c) for c in torch.split(P.reshape(-1, 3), 10000, dim=0)]).reshape(-1, self.n_steps)
        key = torch.sign(vals) * torch.arange(self.n_steps, 0, -1).float().reshape(1, -1)
        first = torch.argmin(key, -1)
        ar = torch.arange(P.shape[0])
        if self.margin is not None:
            # samples whose sign decides `first` / `net_surf`: all up to the first negative one, or all 100
            neg_found = vals[ar, first] < 0
            upto = torch.where(neg_found, first, torch.full_like(first, self.n_steps - 1))
            decisive = torch.arange(self.n_steps).reshape(1, -1) <= upto.reshape(-1, 1)
            mg = torch.where(decisive, vals.abs(), torch.full_like(vals, float("inf"))).min(-1).values
            self.margin[ridx] = torch.minimum(self.margin[ridx], mg)
        out_pts[ridx] = P[ar, first]
        out_d[ridx] = z[ar, first]
        true_surf = object_mask[samp_mask]
        net_surf = vals[ar, first] < 0
        p_out = ~(true_surf & net_surf)
        if p_out.sum() > 0:                                                  # :221-226
            j = torch.argmin(vals[p_out], -1)
            out_pts[ridx[p_out]] = P[p_out][torch.arange(int(p_out.sum())), j]
            out_d[ridx[p_out]] = z[p_out][torch.arange(int(p_out.sum())), j]
        net_mask = samp_mask.clone()
        net_mask[ridx[~net_surf]] = False
        sec = (net_surf & true_surf) if self.training else net_surf          # :233
        n_sec = int(sec.sum())
        self.stats["n_secant"] = n_sec
        if n_sec > 0:
            z_hi = z[ar, first][sec]
            s_hi = vals[ar, first][sec]
            z_lo = z[sec][torch.arange(n_sec), first[sec] - 1]               # index -1 wraps (:239)
            s_lo = vals[sec][torch.arange(n_sec), first[sec] - 1]
            cam_s = cam_loc.unsqueeze(1).repeat(1, N, 1).reshape(-1, 3)[ridx[sec]]
            dir_s = dirs.reshape(-1, 3)[ridx[sec]]
            zp = self._secant(s_lo, s_hi, z_lo, z_hi, cam_s, dir_s, f)
            out_pts[ridx[sec]] = cam_s + zp.unsqueeze(-1) * dir_s
            out_d[ridx[sec]] = zp
        return out_pts, net_mask, out_d

    def _secant(self, s_lo, s_hi, z_lo, z_hi, cam, dirs, f):                 # :251-268
        zp = -s_lo * (z_hi - z_lo) / (s_hi - s_lo) + z_lo
        for _ in range(self.n_secant):
            s_mid = f(cam + zp.unsqueeze(-1) * dirs)
            lo = s_mid > 0
            z_lo[lo] = zp[lo]
            s_lo[lo] = s_mid[lo]
            hi = s_mid < 0
            z_hi[hi] = zp[hi]
            s_hi[hi] = s_mid[hi]
            zp = -s_lo * (z_hi - z_lo) / (s_hi - s_lo) + z_lo
        return zp

    def _min_sdf(self, N, f, cam_loc, dirs, m, min_dis, max_dis, steps_u):   # :270-298
        n_m = int(m.sum())
        n = self.n_steps
        u = steps_u if steps_u is not None else torch.empty(n).uniform_(0.0, 1.0)
        hi = max_dis[m].unsqueeze(-1)
        lo = min_dis[m].unsqueeze(-1)
        steps = u.unsqueeze(0).repeat(n_m, 1) * (hi - lo) + lo
        cam_m = cam_loc.unsqueeze(1).repeat(1, N, 1).reshape(-1, 3)[m]
        dir_m = dirs[m]
        allp = cam_m.unsqueeze(1).repeat(1, n, 1) + steps.unsqueeze(-1) * dir_m.unsqueeze(1).repeat(1, n, 1)
        vals = torch.cat([f(c) for c in torch.split(allp.reshape(-1, 3), 10000, dim=0)]).reshape(-1, n)
        _, j = vals.min(-1)
        ar = torch.arange(n_m)
        return allp.reshape(-1, n, 3)[ar, j], steps[ar, j]


# --------------------------------------------------------------------------------------
# IDRNetwork.forward / loss                   implicit_differentiable_renderer.py:242-329
# --------------------------------------------------------------------------------------
class IDRCfg:
    def __init__(self, embed: Optional[EmbedCfg], n_lin_sdf=9, skip_in=(4,), n_lin_rgb=5,
                 multires_view=4, view_embed: Optional[EmbedCfg] = None, ray_tracer: Optional[dict] = None,
                 bounding_sphere=1.0):
        self.embed = embed
        self.n_lin_sdf = n_lin_sdf
        self.skip_in = tuple(skip_in)
        self.n_lin_rgb = n_lin_rgb
        self.multires_view = multires_view
        self.view_embed = view_embed
        self.ray_tracer = ray_tracer or {}
        self.bounding_sphere = bounding_sphere


def sample_network(surface_output, surface_sdf_values, surface_points_grad, surface_dists,
                   surface_cam_loc, surface_ray_dirs):
    """SampleNetwork.forward (IDR eq. 3).  sample_network.py:10-20."""
    d0 = surface_ray_dirs.detach()
    dot = (surface_points_grad * d0).sum(-1, keepdim=True)
    t = surface_dists - (surface_output - surface_sdf_values) / dot
    return surface_cam_loc + t * surface_ray_dirs


def idr_forward(inp: Dict[str, Tensor], sd: Dict[str, Tensor], cfg: IDRCfg, training: bool = True,
                eik_points: Optional[Tensor] = None, min_sdf_steps: Optional[Tensor] = None,
                tracer_out=None, stats: Optional[dict] = None) -> Dict[str, Tensor]:
    """IDRNetwork.forward.  implicit_differentiable_renderer.py:242-319.
    `eik_points` / `min_sdf_steps` inject the two CPU-RNG draws (:279, ray_tracing.py:277)."""
    uv, pose, K = inp["uv"], inp["pose"], inp["intrinsics"]
    object_mask = inp["object_mask"].reshape(-1)
    dirs, cam_loc = camera_rays(uv, pose, K)
    B, N, _ = dirs.shape

    def net(x):
        return implicit_forward(x, sd, cfg.embed, cfg.n_lin_sdf, cfg.skip_in)

    def grad(x):
        return implicit_gradient(x, sd, cfg.embed, n_lin=cfg.n_lin_sdf, skip_in=cfg.skip_in)

    if tracer_out is None:
        tracer = RayTracerOracle(**cfg.ray_tracer)
        tracer.training = training
        with torch.no_grad():
            points, net_mask, dists = tracer(lambda x: net(x)[:, 0], cam_loc, object_mask, dirs, min_sdf_steps)
        if stats is not None:
            stats.update(tracer.stats)
    else:
        points, net_mask, dists = tracer_out
    points = (cam_loc.unsqueeze(1) + dists.reshape(B, N, 1) * dirs).reshape(-1, 3)
    sdf_output = net(points)[:, 0:1]
    dirs = dirs.reshape(-1, 3)
    if training:
        surf = net_mask & object_mask
        sp = points[surf]
        sdist = dists[surf].unsqueeze(-1)
        sdirs = dirs[surf]
        scam = cam_loc.unsqueeze(1).repeat(1, N, 1).reshape(-1, 3)[surf]
        sout = sdf_output[surf]
        n_s = sp.shape[0]
        if eik_points is None:
            eik_points = torch.empty(B * N // 2, 3).uniform_(-cfg.bounding_sphere, cfg.bounding_sphere)
        eik = torch.cat([eik_points, points.clone().detach()], 0)
        allp = torch.cat([sp, eik], dim=0)
        s0 = net(sp)[:n_s, 0:1].detach()
        g = grad(allp)
        sgrad = g[:n_s, 0, :].clone().detach()
        grad_theta = g[n_s:, 0, :]
        diff_pts = sample_network(sout, s0, sgrad, sdist, scam, sdirs)
    else:
        surf = net_mask
        diff_pts = points[surf]
        grad_theta = None
    view = -dirs[surf]
    rgb = torch.ones_like(points)            # the reference adds .float(): points are fp32 there
    if diff_pts.shape[0] > 0:                                               # get_rbg_value :321-329
        out = net(diff_pts)
        nrm = grad(diff_pts)[:, 0, :]
        rgb[surf] = rendering_forward(diff_pts, nrm, view, out[:, 1:], sd, cfg.multires_view,
                                      cfg.n_lin_rgb, cfg.view_embed)
    return {"points": points, "rgb_values": rgb, "sdf_output": sdf_output,
            "network_object_mask": net_mask, "object_mask": object_mask, "grad_theta": grad_theta}


def idr_loss(out: Dict[str, Tensor], rgb_gt: Tensor, eikonal_weight=0.1, mask_weight=100.0,
             alpha=50.0) -> Dict[str, Tensor]:
    """IDRLoss.forward.  loss.py:5-71."""
    net_mask, obj_mask = out["network_object_mask"], out["object_mask"]
    both = net_mask & obj_mask
    n = float(obj_mask.shape[0])
    if both.sum() == 0:
        rgb_loss = torch.tensor(0.0)
    else:
        rgb_loss = (out["rgb_values"][both] - rgb_gt.reshape(-1, 3)[both]).abs().sum() / n
    gt = out["grad_theta"]
    eik = torch.tensor(0.0) if gt.shape[0] == 0 else ((gt.norm(2, dim=1) - 1) ** 2).mean()
    neg = ~both
    if neg.sum() == 0:
        mask_loss = torch.tensor(0.0)
    else:
        logits = -alpha * out["sdf_output"][neg]
        mask_loss = (1 / alpha) * F.binary_cross_entropy_with_logits(
            logits.squeeze(-1), obj_mask[neg].float(), reduction="sum") / n
    loss = rgb_loss + eikonal_weight * eik + mask_weight * mask_loss
    return {"loss": loss, "rgb_loss": rgb_loss, "eikonal_loss": eik, "mask_loss": mask_loss}


# --------------------------------------------------------------------------------------
# Parameter construction (reference-shaped state dicts with our own RNG)
# --------------------------------------------------------------------------------------
def make_hashgrid_sd(prefix: str, n_levels: int, n_feat: int, log2_T: int, base: int, desired: int,
                     gen: torch.Generator, table_std: float = 1e-4) -> Dict[str, Tensor]:
    """Random tables U(-1e-4,1e-4) (hashGridEmbedding.py:69-71) and B ~ N(0, sigma^2) (:141)."""
    _, rows = level_schedule(n_levels, base, desired, log2_T)
    sd = {}
    for l, t in enumerate(rows):
        sd[f"{prefix}levels.{l}.embedding.weight"] = (torch.rand(t, n_feat, generator=gen) * 2 - 1) * table_std
    sd[f"{prefix}freq_encoding.B"] = torch.randn(3, n_levels, generator=gen) * fourier_sigma(base, desired)
    return sd


def make_nffb_sd(prefix: str, n_levels: int, n_feat: int, log2_T: int, base: int, desired: int,
                 style: bool, gen: torch.Generator) -> Dict[str, Tensor]:
    """SIREN init of the filter-bank trunk.  nffb3d.py:77-79,114,237-243; Sine.py:14-25."""
    sd = make_hashgrid_sd(f"{prefix}grid_enc.", n_levels, n_feat, log2_T, base, desired, gen)
    W = 2 * (n_feat * (2 + 2 * n_levels))
    w0 = float(n_levels ** n_feat - n_levels)

    def uni(shape, a):
        return (torch.rand(*shape, generator=gen) * 2 - 1) * a

    sd[f"{prefix}ff_lin0.weight"] = uni((W, 3), 1.0 / 3)
    sd[f"{prefix}ff_lin0.bias"] = uni((W,), 1.0 / 3)
    for j in range(1, n_levels - 1):
        a = math.sqrt(6.0 / W) / w0
        sd[f"{prefix}ff_lin{j}.weight"] = uni((W, W), a)
        sd[f"{prefix}ff_lin{j}.bias"] = uni((W,), a)
    a = 1.0 / math.sqrt(W)
    sd[f"{prefix}out_layer.weight"] = uni((W, W), a)
    sd[f"{prefix}out_layer.bias"] = uni((W,), a)
    if style:
        sd[f"{prefix}StyleAttentionBlock.linear_transform.weight"] = uni((W, W), a)
        sd[f"{prefix}StyleAttentionBlock.linear_transform.bias"] = uni((W,), a)
        sd[f"{prefix}StyleAttentionBlock.attention.weight"] = uni((1, 3), 1.0 / math.sqrt(3))
        sd[f"{prefix}StyleAttentionBlock.attention.bias"] = uni((1,), 1.0 / math.sqrt(3))
    return sd


def make_embed_sd(prefix: str, cfg: EmbedCfg, gen: torch.Generator) -> Dict[str, Tensor]:
    pre = f"{prefix}embedder_obj."
    if cfg.embed_type == "HashGrid":
        return make_hashgrid_sd(pre, cfg.multires, cfg.n_feat, cfg.log2_T, cfg.base, cfg.desired, gen)
    if cfg.embed_type in ("FFB", "StyleModNFFB"):
        return make_nffb_sd(pre, cfg.multires, cfg.n_feat, cfg.log2_T, cfg.base, cfg.desired,
                            cfg.embed_type == "StyleModNFFB", gen)
    if cfg.embed_type == "FourierFeatures":
        return {f"{pre}B": torch.randn(3, 3, generator=gen)}
    return {}


def make_implicit_sd(cfg: Optional[EmbedCfg], gen: torch.Generator, feature: int = 256,
                     dims: Sequence[int] = (512,) * 8, skip_in=(4,), bias: float = 0.6,
                     prefix: str = "implicit_network.", perturb: float = 0.0) -> Dict[str, Tensor]:
    """Geometric initialisation (implicit_differentiable_renderer.py:54-81) expressed as
    (weight_g, weight_v, bias) triples.  `perturb` adds noise so outputs are not degenerate."""
    sd = make_embed_sd(f"{prefix}embed_model.", cfg, gen) if cfg is not None else {}
    d0 = cfg.width() if cfg is not None else 3
    full = [d0] + list(dims) + [1 + feature]
    n = len(full)
    for l in range(n - 1):
        out_dim = full[l + 1] - full[0] if (l + 1) in skip_in else full[l + 1]
        w = torch.zeros(out_dim, full[l])
        b = torch.zeros(out_dim)
        if l == n - 2:
            w = torch.randn(out_dim, full[l], generator=gen) * 1e-4 + math.sqrt(math.pi) / math.sqrt(full[l])
            b = b - bias
        elif l == 0 and cfg is not None:
            w[:, :3] = torch.randn(out_dim, 3, generator=gen) * (math.sqrt(2) / math.sqrt(out_dim))
        elif l in skip_in and cfg is not None:
            w = torch.randn(out_dim, full[l], generator=gen) * (math.sqrt(2) / math.sqrt(out_dim))
            w[:, -(full[0] - 3):] = 0.0
        else:
            w = torch.randn(out_dim, full[l], generator=gen) * (math.sqrt(2) / math.sqrt(out_dim))
        if perturb > 0:
            w = w + torch.randn(w.shape, generator=gen) * perturb
            b = b + torch.randn(b.shape, generator=gen) * perturb
        sd[f"{prefix}lin{l}.weight_g"] = w.norm(2, dim=1, keepdim=True)
        sd[f"{prefix}lin{l}.weight_v"] = w
        sd[f"{prefix}lin{l}.bias"] = b
    sd[f"{prefix}dencity_net.beta"] = torch.tensor(0.9)
    return sd


def make_rendering_sd(gen: torch.Generator, d_in0: int, dims=(512,) * 4, d_out: int = 3,
                      prefix: str = "rendering_network.") -> Dict[str, Tensor]:
    full = [d_in0] + list(dims) + [d_out]
    sd = {}
    for l in range(len(full) - 1):
        a = 1.0 / math.sqrt(full[l])
        w = (torch.rand(full[l + 1], full[l], generator=gen) * 2 - 1) * a
        sd[f"{prefix}lin{l}.weight_g"] = w.norm(2, dim=1, keepdim=True)
        sd[f"{prefix}lin{l}.weight_v"] = w
        sd[f"{prefix}lin{l}.bias"] = (torch.rand(full[l + 1], generator=gen) * 2 - 1) * a
    return sd


def make_idr_sd(cfg: IDRCfg, seed: int = 0, perturb: float = 0.0) -> Dict[str, Tensor]:
    gen = torch.Generator().manual_seed(seed)
    sd = make_implicit_sd(cfg.embed, gen, skip_in=cfg.skip_in, perturb=perturb)
    if cfg.view_embed is not None:
        sd.update(make_embed_sd("rendering_network.embed_model.", cfg.view_embed, gen))
        d0 = 9 + 256 + (cfg.view_embed.width() - 3)
    else:
        d0 = 9 + 256 + (3 * (1 + 2 * cfg.multires_view) if cfg.multires_view > 0 else 0)
    sd.update(make_rendering_sd(gen, d0))
    return sd


def synthetic_batch(n_rays: int, seed: int = 1) -> Tuple[Dict[str, Tensor], Tensor]:
    """SURVEY.md §8(d) synthetic camera: pose = I, t = (0,0,-3), f = 500, c = 128, uv in [0,256)."""
    pose = torch.eye(4).unsqueeze(0)
    pose[0, 2, 3] = -3.0
    K = torch.eye(4).unsqueeze(0)
    K[0, 0, 0] = K[0, 1, 1] = 500.0
    K[0, 0, 2] = K[0, 1, 2] = 128.0
    uv = torch.rand(1, n_rays, 2, generator=torch.Generator().manual_seed(seed)) * 256
    mask = torch.rand(1, n_rays, generator=torch.Generator().manual_seed(seed + 1)) > 0.5
    rgb = torch.rand(1, n_rays, 3, generator=torch.Generator().manual_seed(seed + 2)) * 2 - 1
    return {"uv": uv, "pose": pose, "intrinsics": K, "object_mask": mask}, rgb
