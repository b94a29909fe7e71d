"""StyleAttention with the reference's parameters (style_Attention/styleMod.py:17-44).

What the reference computes: softmax over a size-1 dimension is identically 1, and InstanceNorm1d
applied to a 2-D tensor normalises every row over its features (biased variance, eps 1e-5, no affine).
So the block is  rownorm(linear_transform(style));  `attention` (Linear d_in -> 1) is kept for
state-dict compatibility and never influences the output."""
import torch
import torch.nn as nn

from .... import mlp


class StyleAttention(nn.Module):
    def __init__(self, d_in=3, feature_vector_size=28):
        super().__init__()
        self.d_in = d_in
        self.feature_vector_size = feature_vector_size
        self.linear_transform = nn.Linear(feature_vector_size, feature_vector_size)
        self.attention = nn.Linear(d_in, 1)
        self.eps = 1e-5

    def forward(self, content, style):
        style = style.reshape(-1, self.feature_vector_size)
        y = mlp.linear(style, self.linear_transform.weight, self.linear_transform.bias)
        # per-row normalisation over the features, biased variance, no affine = layer_norm over the last dimension: one
        # fused forward and one fused backward kernel (differentiable again for the recorded backward) instead of the
        # mean / var / sub / sqrt / div chain and its ~15 backward ops
        return torch.nn.functional.layer_norm(y, (self.feature_vector_size,), None, None, self.eps)
