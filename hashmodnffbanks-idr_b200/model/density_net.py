"""Laplace density used to squash the SDF channel (reference model/density_net.py:16-30).
`beta` is a parameter for state-dict compatibility (`dencity_net.beta`); the density is always
evaluated without gradient, exactly as the reference's @torch.no_grad() density_func."""
import torch
import torch.nn as nn


class Density(nn.Module):
    def __init__(self, params_init={}):
        super().__init__()
        for name, value in params_init.items():
            setattr(self, name, nn.Parameter(torch.tensor(value)))

    def forward(self, sdf, beta=None, compute_grad=False):
        return self.density_func(sdf, beta=beta)


class LaplaceDensity(Density):
    """(1/beta) * (0.5 + 0.5 * sign(s) * expm1(-|s| / beta)),  beta = |beta_param| + beta_min."""

    def __init__(self, params_init={}, beta_min=0.0001):
        super().__init__(params_init=params_init)
        self.beta_min = torch.tensor(beta_min)

    def get_beta(self):
        return self.beta.abs() + float(self.beta_min)

    @torch.no_grad()
    def density_func(self, sdf, beta=None):
        b = self.get_beta()
        return (1.0 / b) * (0.5 + 0.5 * sdf.sign() * torch.expm1(-sdf.abs() / b))
