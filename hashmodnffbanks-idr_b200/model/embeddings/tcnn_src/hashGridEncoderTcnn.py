"""`MultiResHashGridEncoderTcnn` with the reference's constructor and attributes
(model/embeddings/tcnn_src/hashGridEncoderTcnn.py:8-88) - WITHOUT tiny-cuda-nn.

The reference builds `tcnn.Encoding(n_input_dims, {"otype": "Grid", "type": "Hash", "interpolation": "Linear", ...})`
(:63-80) from an un-vendored, unpinned dependency.  Here the same grid is one launch of this repository's hash-encode
kernel in its IDRK_HASH_NGP mode (include/idrk.h), following tiny-cuda-nn's published algorithm:

  * level l: scale_l = base_resolution * per_level_scale**l - 1 (fp32), resolution R_l = ceil(scale_l) + 1;
  * parameters per level = min(next_multiple(R_l^3, 8), 2**log2_hashmap_size) rows of F features, all levels in ONE
    flat fp32 vector `grid_encoder.params` (tcnn's layout and state-dict key), initialised U(-1e-4, 1e-4);
  * a point x in [0, 1]^3 sits at pos = x * scale_l + 0.5; its 8 surrounding vertices are weighted trilinearly;
  * vertex -> row: x + y R + z R^2 while the level fits its table (dense), else the coherent prime hash
    x ^ (y * 2654435761) ^ (z * 805459861); mod rows.

Parity is UNPINNED for this module (tiny-cuda-nn is absent here and the reference ships no vectors
for it); tests pin the kernel to oracle/idr_oracle.py::ngp_grid_encode and to interpolation identities instead.
Differences on purpose: fp32 tables and outputs (tcnn's torch binding defaults to fp16), keys tcnn would ignore
(`hidden_dims`, `base_sigma`, `exp_sigma`, `grid_embedding_std`) are accepted and unused like in the reference call.
"""
import math

import numpy as np
import torch
import torch.nn as nn

from .... import autograd_ops as ops
from .... import kernels as K
from ...._lib import HASH_NGP


def ngp_level_layout(n_levels, n_features, log2_hashmap_size, base_resolution, per_level_scale):
    """(scales, resolutions, rows, offsets) of tcnn's Hash grid; offsets in rows, len L + 1."""
    log2_pls = np.float32(math.log2(per_level_scale))
    scales, res, rows, offs = [], [], [], [0]
    for l in range(n_levels):
        scale = np.float32(np.exp2(np.float32(l) * log2_pls) * np.float32(base_resolution) - np.float32(1.0))
        R = int(math.ceil(float(scale))) + 1
        n = min(R ** 3, (2 ** 32 - 1) // 2)
        n = (n + 7) // 8 * 8
        n = min(n, 1 << log2_hashmap_size)
        scales.append(float(scale)); res.append(R); rows.append(n); offs.append(offs[-1] + n)
    return scales, res, rows, offs


class NgpGrid(nn.Module):
    """Stands where `tcnn.Encoding` stands in the reference: `.params` is the flat parameter vector, calling it on
    x [n, 3] in [0, 1] returns the [n, L * F] level features."""

    def __init__(self, n_levels, n_features, log2_hashmap_size, base_resolution, per_level_scale):
        super().__init__()
        if n_features != 2:
            raise ValueError("idrk NGP grid: n_features_per_level must be 2 (the value every reference config uses)")
        self.n_levels, self.n_features = int(n_levels), int(n_features)
        self.scales, self.resolutions, self.rows, self.offsets = ngp_level_layout(
            n_levels, n_features, log2_hashmap_size, base_resolution, per_level_scale)
        self.n_output_dims = self.n_levels * self.n_features
        self.params = nn.Parameter(torch.empty(self.offsets[-1] * self.n_features).uniform_(-1e-4, 1e-4))
        self._spec = K.HashGridSpec(self.scales, self.rows, self.n_features, HASH_NGP, 0)

    def spec(self):
        return self._spec

    def tables(self):
        F = self.n_features
        return tuple(self.params[self.offsets[l] * F: self.offsets[l + 1] * F].view(self.rows[l], F)
                     for l in range(self.n_levels))

    def forward(self, x):
        return ops.hash_encode(x, self._spec, self.tables(), None)


class MultiResHashGridEncoderTcnn(nn.Module):
    def __init__(self, include_input: bool, in_dim: int, network_dims: list, embed_type: str, n_levels: int,
                 max_points_per_level: int, log2_hashmap_size: int, base_resolution: int, desired_resolution: int,
                 base_sigma: float, exp_sigma: float, grid_embedding_std: float, per_level_scale: float):
        super().__init__()
        if in_dim != 3:
            raise ValueError("idrk hash grid supports 3-D inputs")
        if embed_type != 'HashGridTcnn':
            raise ValueError("embed_type must be 'HashGridTcnn'")          # the reference leaves otype unbound otherwise
        self.in_dim = in_dim
        self.include_input = include_input
        self.grid_encoder = NgpGrid(int(n_levels), max_points_per_level, log2_hashmap_size, base_resolution,
                                    per_level_scale)
        self.grid_levels = n_levels
        self.output_dim = self.grid_levels * max_points_per_level
        self.embeddings_dim = self.in_dim + self.output_dim if include_input else self.output_dim

    def forward(self, x, compute_grad=False):
        g = self.grid_encoder(x)
        return torch.cat([x, g], dim=-1) if self.include_input else g
