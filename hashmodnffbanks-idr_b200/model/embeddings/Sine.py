"""SIREN activation and initialisers with the reference's names (model/embeddings/Sine.py:5-25).
Inside FourierFilterBanks the activation is fused into the contraction kernel's epilogue."""
import numpy as np
import torch
from torch import nn


class Sine(nn.Module):
    def __init__(self, w0):
        super().__init__()
        self.w0 = w0

    def forward(self, input, compute_grad=False):
        return torch.sin(input * self.w0)


def sine_init(m, w0, num_input=None):
    if hasattr(m, 'weight') and num_input is None:
        fan_in = m.weight.size(-1)
        bound = np.sqrt(6 / fan_in) / w0
        nn.init.uniform_(m.weight, -bound, bound)
        nn.init.uniform_(m.bias, -bound, bound)


def first_layer_sine_init(m):
    if hasattr(m, 'weight'):
        fan_in = m.weight.size(-1)
        nn.init.uniform_(m.weight, -1.0 / fan_in, 1.0 / fan_in)
        nn.init.uniform_(m.bias, -1.0 / fan_in, 1.0 / fan_in)
