"""Multi-resolution hash grid with the reference's module interface
(model/embeddings/hashGridEmbedding.py): `_HashGridMLP` (:42-102) holds one level's table as
`embedding.weight`, `MultiResHashGridMLP` (:105-155) owns `levels` and `freq_encoding`, so the
state_dict keys are `levels.{l}.embedding.weight` and `freq_encoding.B`.

All levels and the Fourier prefix are evaluated by ONE kernel launch (csrc/hash_encode.cu).
`frac_mode="reference"` (default) reproduces the reference bit-exactly - its fractional part is
identically zero (hashGridEmbedding.py:86), so a level returns its floor-corner row;
`frac_mode="trilinear"` interpolates the 8 corners (documented extension, not the parity target).
"""
import math

import torch
import torch.nn as nn

from ... import autograd_ops as ops
from ... import kernels as K
from ..._lib import HASH_REFERENCE, HASH_TRILINEAR
from .frequency_enc import FourierFeature as FrequencyEncoding

HASH_PRIMES = [1, 3, 2654435761]          # the three primes the 3-D path uses (reference :14)

_MODES = {"reference": HASH_REFERENCE, "trilinear": HASH_TRILINEAR}


def level_resolutions(n_levels, base_resolution, desired_resolution):
    """floor(base * growth**l) with growth from Python doubles, as the reference ctor (:125-130)."""
    growth = math.exp((math.log(desired_resolution) - math.log(base_resolution)) / (n_levels - 1))
    return [math.floor(base_resolution * (growth ** l)) for l in range(n_levels)]


class _HashGridMLP(nn.Module):
    """One resolution level: a [hashmap_size, n_features] table initialised U(-1e-4, 1e-4)."""

    def __init__(self, dim: int, n_features: int, hashmap_size: int, resolution: float, frac_mode: str = "reference"):
        super().__init__()
        if dim != 3:
            raise ValueError("idrk hash grid supports 3-D inputs")
        self.dim, self.n_features = dim, n_features
        self.hashmap_size, self.resolution = int(hashmap_size), resolution
        self.frac_mode = frac_mode
        self.embedding = nn.Embedding(self.hashmap_size, n_features)
        nn.init.uniform_(self.embedding.weight, -1e-4, 1e-4)

    def forward(self, x: torch.Tensor, compute_grad=False) -> torch.Tensor:
        spec = K.HashGridSpec([self.resolution], [self.hashmap_size], self.n_features, _MODES[self.frac_mode], 0)
        return ops.hash_encode(x, spec, (self.embedding.weight,), None)


class MultiResHashGridMLP(nn.Module):
    def __init__(self, include_input: bool, in_dim: int, n_levels: int, max_points_per_level: int,
                 log2_hashmap_size: int, base_resolution: int, desired_resolution: int,
                 frac_mode: str = "reference"):
        super().__init__()
        if frac_mode not in _MODES:
            raise ValueError("frac_mode must be 'reference' or 'trilinear'")
        self.include_input = include_input
        self.frac_mode = frac_mode
        self.n_levels, self.n_features = n_levels, max_points_per_level
        res = level_resolutions(n_levels, base_resolution, desired_resolution)
        levels = []
        for r in res:
            self.hashmap_size = min(r ** in_dim, 2 ** log2_hashmap_size)
            levels.append(_HashGridMLP(in_dim, max_points_per_level, self.hashmap_size, r, frac_mode))
        self.levels = nn.Sequential(*levels)
        if include_input:
            sigma = (math.log(desired_resolution) - math.log(base_resolution)) / (base_resolution - 1)
            self.freq_encoding = FrequencyEncoding(in_dim, sigma, num_channels=n_levels, include_input=True)
            self.embeddings_dim = in_dim + n_levels * max_points_per_level + (self.freq_encoding.embeddings_dim - in_dim)
        else:
            self.embeddings_dim = n_levels * max_points_per_level
        self._spec = K.HashGridSpec(res, [l.hashmap_size for l in levels], max_points_per_level,
                                    _MODES[frac_mode], n_levels if include_input else 0)

    def tables(self):
        return tuple(l.embedding.weight for l in self.levels)

    def spec(self) -> K.HashGridSpec:
        return self._spec

    def forward(self, x: torch.Tensor, compute_grad=False) -> torch.Tensor:
        B = self.freq_encoding.B if self.include_input else None
        return ops.hash_encode(x, self._spec, self.tables(), B)

    @torch.no_grad()
    def corner_indices(self, x: torch.Tensor) -> torch.Tensor:
        """uint32 table row of the 8 corners for every (point, level): [n, L, 8] (parity/debug)."""
        B = self.freq_encoding.B if self.include_input else None
        _, idx = K.hash_encode_fwd(self._spec, x.reshape(-1, 3), self.tables(), B, want_idx=True)
        return idx
