"""Frequency encoders with the reference's class names and constructor arguments
(model/embeddings/frequency_enc.py): PositionalEncoding (:6-51), FourierFeature (:54-67),
SHEncoder (:70-152), get_embedder (:156-168).  Arithmetic runs in libidrk kernels."""
import numpy as np
import torch
import torch.nn as nn

from ... import autograd_ops as ops


class PositionalEncoding(nn.Module):
    """NeRF positional encoding.  Emits [x, x, sin(f0 x), cos(f0 x), ...] when include_input is
    set - the input appears twice, exactly like the reference (frequency_enc.py:24,46-47)."""

    def __init__(self, **kwargs):
        super().__init__()
        self.kwargs = kwargs
        self.include_input = bool(kwargs['include_input'])
        d = int(kwargs['input_dims'])
        n_freqs = int(kwargs['num_freqs'])
        max_freq = kwargs['max_freq_log2']
        if kwargs.get('log_sampling', True):
            bands = 2. ** torch.linspace(0., max_freq, n_freqs)
        else:
            bands = torch.linspace(2. ** 0., 2. ** max_freq, n_freqs)
        fns = kwargs.get('periodic_fns', [torch.sin, torch.cos])
        if list(fns) != [torch.sin, torch.cos]:
            raise ValueError("idrk PositionalEncoding supports periodic_fns=[torch.sin, torch.cos] only")
        self.freq_bands = [float(b) for b in bands]
        # the reference reports widths computed from `input_dims` even when wider rows are fed in
        self.out_dim = d + 2 * d * n_freqs
        self.embeddings_dim = self.out_dim + d if self.include_input else self.out_dim

    def embed(self, inputs):
        return ops.positional_encoding(inputs, self.freq_bands, self.include_input)

    def forward(self, inputs, compute_grad=False):
        return self.embed(inputs)


class FourierFeature(nn.Module):
    """Random Fourier features [x | sin(2 pi x B) | cos(2 pi x B)], B a persistent buffer."""

    def __init__(self, input_dims=3, sigma=1.0, num_channels=256, include_input=True) -> None:
        super().__init__()
        if input_dims != 3:
            raise ValueError("idrk FourierFeature supports input_dims=3")
        self.input_dims = input_dims
        self.include_input = include_input
        self.register_buffer('B', torch.randn(input_dims, int(num_channels)) * sigma, persistent=True)
        self.embeddings_dim = 2 * num_channels + 3 if include_input else 2 * num_channels

    def forward(self, x, compute_grad=False):
        full = ops.fourier_feature(x, self.B)
        return full if self.include_input else full[..., 3:]


class SHEncoder(nn.Module):
    """Spherical-harmonics view encoder (optional `viewdirs_embed_type = SHEncoder`).  Low-degree
    polynomial of the unit direction; evaluated with elementwise device ops (not a hot-path kernel)."""

    def __init__(self, input_dims=3, degree=4):
        super().__init__()
        assert input_dims == 3 and 1 <= degree <= 5
        self.input_dims, self.degree = input_dims, degree
        self.embeddings_dim = degree ** 2

    def forward(self, input, **kwargs):
        x, y, z = input.unbind(-1)
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        c = [torch.full_like(x, 0.28209479177387814)]
        if self.degree > 1:
            c += [-0.4886025119029199 * y, 0.4886025119029199 * z, -0.4886025119029199 * x]
        if self.degree > 2:
            c += [1.0925484305920792 * xy, -1.0925484305920792 * yz, 0.31539156525252005 * (2.0 * zz - xx - yy),
                  -1.0925484305920792 * xz, 0.5462742152960396 * (xx - yy)]
        if self.degree > 3:
            c += [-0.5900435899266435 * y * (3 * xx - yy), 2.890611442640554 * xy * z,
                  -0.4570457994644658 * y * (4 * zz - xx - yy), 0.3731763325901154 * z * (2 * zz - 3 * xx - 3 * yy),
                  -0.4570457994644658 * x * (4 * zz - xx - yy), 1.445305721320277 * z * (xx - yy),
                  -0.5900435899266435 * x * (xx - 3 * yy)]
        if self.degree > 4:
            c += [2.5033429417967046 * xy * (xx - yy), -1.7701307697799304 * yz * (3 * xx - yy),
                  0.9461746957575601 * xy * (7 * zz - 1), -0.6690465435572892 * yz * (7 * zz - 3),
                  0.10578554691520431 * (zz * (35 * zz - 30) + 3), -0.6690465435572892 * xz * (7 * zz - 3),
                  0.47308734787878004 * (xx - yy) * (7 * zz - 1), -1.7701307697799304 * xz * (xx - 3 * yy),
                  0.6258357354491761 * (xx * (xx - 3 * yy) - yy * (3 * xx - yy))]
        return torch.stack(c, dim=-1)


def get_embedder(multires):
    """Default IDR view-direction embedder: returns (embed_fn, out_dim) with out_dim = 3(1+2*multires)
    although embed_fn emits 3(2+2*multires) columns (frequency_enc.py:168)."""
    eo = PositionalEncoding(include_input=True, input_dims=3, max_freq_log2=multires - 1, num_freqs=multires,
                            log_sampling=True, periodic_fns=[torch.sin, torch.cos])

    def embed(x, eo=eo):
        return eo.embed(x)
    return embed, eo.out_dim
