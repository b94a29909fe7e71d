"""Fourier filter banks (NFFB / StyleModNFFB) with the reference's constructor, attributes and
state-dict keys (model/embeddings/nffb3d.py:26-194).

Data flow of the live configuration (freq_enc_type='PositionalEncodingNET', layers_type='SIREN',
has_out=False), restated from nffb3d.py:122-194:
    z0 = p / bound,  u = (p + bound) / (2 bound)
    g  = grid_enc(u)[:, 3:]  viewed as L chunks of 2F columns   (chunks mix Fourier and hash columns)
    E_i = PositionalEncoding(chunk_i)                             (width W = 2F(2 + 2L) ... = 2 * posenc dim)
    z_{j+1} = sin(w0 (A_j z_j + a_j)),  w0 = L**F - L
    for j >= 1:  e = (style ? rownorm(M E_{j-1} + m) : E_{j-1}) + z_{j+1};   f += O e + o
    out = [u | f / L]
Only chunks 0..L-3 are consumed, so the others are not encoded here.
Every Linear (+ sine) runs in the contraction kernel, the grid in the hash-encode kernel, the per-chunk
encodings in the posenc kernel; all of them are twice differentiable (see idrk.mlp / idrk.autograd_ops).
"""
import torch
import torch.nn as nn

from ... import autograd_ops as ops
from ... import mlp
from .frequency_enc import PositionalEncoding
from .hashGridEmbedding import MultiResHashGridMLP
from .Sine import Sine, first_layer_sine_init, sine_init
from .style_Attention.styleMod import StyleAttention


class FourierFilterBanks(nn.Module):
    WIDTH_MULT = 2          # trunk width = 2 x the positional encoding's width (nffb3d.py:67-69)

    def _make_grid(self, cfg):
        return MultiResHashGridMLP(self.include_input, self.num_inputs, self.n_levels, self.max_points_per_level,
                                   cfg['log2_hashmap_size'], cfg['base_resolution'], cfg['desired_resolution'])

    def _chunk_width(self):
        return 2 * self.max_points_per_level          # the reference's chunking of the grid columns (nffb3d.py:137-139)

    def __init__(self, GridEncoderNetConfig, freq_enc_type, has_out, bound, layers_type, style_modulation=False):
        super().__init__()
        cfg = GridEncoderNetConfig
        if freq_enc_type != 'PositionalEncodingNET' or layers_type != 'SIREN' or has_out:
            raise ValueError("idrk FourierFilterBanks implements the configuration the reference's selector uses: "
                             "PositionalEncodingNET + SIREN + has_out=False")
        self.bound = bound
        self.include_input = cfg['include_input']
        self.num_inputs = cfg['in_dim']
        self.n_levels = int(cfg['n_levels'])
        self.max_points_per_level = cfg['max_points_per_level']
        self.network_dims = cfg.get('network_dims')
        self.modulationApplied = style_modulation
        self.grid_levels = self.n_levels
        if 2 + self.max_points_per_level != 2 * self.max_points_per_level:
            raise ValueError("the reference's chunking (nffb3d.py:137-139) requires max_points_per_entry == 2")
        self.grid_enc = self._make_grid(cfg)
        enc = [PositionalEncoding(include_input=self.include_input, input_dims=self.max_points_per_level,
                                  max_freq_log2=self.n_levels - 1, num_freqs=self.n_levels, log_sampling=True,
                                  periodic_fns=[torch.sin, torch.cos]) for _ in range(self.grid_levels)]
        self.ff_enc = nn.Sequential(*enc)
        width = self.WIDTH_MULT * enc[-1].embeddings_dim
        self.nffb_lin_dims = [self.num_inputs] + [width] * (self.grid_levels - 1)
        self.n_nffb_layers = len(self.nffb_lin_dims)
        assert self.n_nffb_layers >= 3, "The NFFB should have at least 3 levels"
        for layer in range(self.n_nffb_layers - 1):
            setattr(self, "ff_lin" + str(layer), nn.Linear(self.nffb_lin_dims[layer], self.nffb_lin_dims[layer + 1]))
        self.sin_w0 = self.n_levels ** self.max_points_per_level - self.n_levels
        self.sin_w0_high = self.sin_w0 + 10
        self.sin_activation = Sine(w0=self.sin_w0)
        self.sin_activation_high = Sine(w0=self.sin_w0_high)
        self.lin_activation = self.sin_activation
        self.init_SIREN()
        self.feature_Vector_size = width
        self.has_out = has_out
        self.embeddings_dim = width + self.num_inputs if self.include_input else width
        self.out_layer = nn.Linear(width, width)
        if self.modulationApplied:
            self.StyleAttentionBlock = StyleAttention(self.num_inputs, self.feature_Vector_size)

    def init_SIREN(self):
        for layer in range(self.n_nffb_layers - 1):
            lin = getattr(self, "ff_lin" + str(layer))
            if layer == 0:
                first_layer_sine_init(lin)
            else:
                sine_init(lin, self.sin_w0)

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        L, F2 = self.grid_levels, self._chunk_width()
        z = input / self.bound
        u = (input + self.bound) / (2 * self.bound)
        grid = self.grid_enc(u)[..., input.shape[-1]:]
        bands = self.ff_enc[0].freq_bands
        # The reference applies out_layer to (e_j + z_{j+1}) of every level and sums the results (nffb3d.py:163-190); the
        # layer is linear, so the sum is taken first and out_layer runs ONCE:  sum_j (O u_j + o) = O (sum_j u_j) + n o.
        # Same value up to the fp32 summation order; 4 of 5 contractions (and their backward / recorded-backward
        # launches) disappear.  csrc/nffb.cu does the same.
        usum, n_out = None, 0
        for layer in range(self.n_nffb_layers - 1):
            lin = getattr(self, "ff_lin" + str(layer))
            z = mlp.linear_act(z, lin.weight, lin.bias, "sine", float(self.sin_w0))
            if layer > 0:
                chunk = grid[:, (layer - 1) * F2: layer * F2]
                e = ops.positional_encoding(chunk, bands, self.include_input)
                if self.modulationApplied:
                    e = self.StyleAttentionBlock(u, e)
                usum = e + z if usum is None else usum + (e + z)
                n_out += 1
        total = mlp.linear(usum, self.out_layer.weight, self.out_layer.bias * float(n_out))
        return torch.cat([u, total / L], dim=-1)
