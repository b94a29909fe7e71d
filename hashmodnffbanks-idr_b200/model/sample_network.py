"""Differentiable ray/surface intersection (IDR eq. 3) - reference model/sample_network.py:10-20.
A handful of [N_s, 1..3] elementwise operations; they stay on the autograd tape as device tensor ops."""
import torch
import torch.nn as nn


class SampleNetwork(nn.Module):
    def forward(self, surface_output, surface_sdf_values, surface_points_grad, surface_dists, surface_cam_loc,
                surface_ray_dirs):
        dirs0 = surface_ray_dirs.detach()
        denom = (surface_points_grad * dirs0).sum(dim=-1, keepdim=True)
        t_theta = surface_dists - (surface_output - surface_sdf_values) / denom
        return surface_cam_loc + t_theta * surface_ray_dirs
