"""IDR loss with the reference's interface (model/loss.py:5-71): masked L1 on rgb, eikonal term and
the alpha-scaled mask BCE, all normalised by the number of rays.  Scalar reductions over [N] / [N,3]
tensors (negligible work, kept as device tensor ops so they sit on the autograd tape)."""
import torch
from torch import nn
from torch.nn import functional as F


class IDRLoss(nn.Module):
    def __init__(self, eikonal_weight, mask_weight, alpha):
        super().__init__()
        self.eikonal_weight = eikonal_weight
        self.mask_weight = mask_weight
        self.alpha = alpha

    # All three terms are written as masked reductions over fixed-shape tensors (no boolean indexing,
    # no data-dependent branches): same values as the reference's indexed sums - an empty selection simply
    # contributes 0 - and the whole loss can be captured in a CUDA graph.
    def get_rgb_loss(self, rgb_values, rgb_gt, network_object_mask, object_mask):
        both = (network_object_mask & object_mask).unsqueeze(-1)
        diff = torch.where(both, rgb_values - rgb_gt.reshape(-1, 3), torch.zeros_like(rgb_values))
        return diff.abs().sum() / float(object_mask.shape[0])

    def get_eikonal_loss(self, grad_theta):
        if grad_theta.shape[0] == 0:
            return torch.zeros((), device=grad_theta.device)
        return ((grad_theta.norm(2, dim=1) - 1) ** 2).mean()

    def get_mask_loss(self, sdf_output, network_object_mask, object_mask):
        neg = ~(network_object_mask & object_mask)
        logits = -self.alpha * sdf_output.reshape(-1)
        bce = F.binary_cross_entropy_with_logits(logits, object_mask.float(), reduction='none')
        total = torch.where(neg, bce, torch.zeros_like(bce)).sum()
        return (1 / self.alpha) * total / float(object_mask.shape[0])

    def forward(self, model_outputs, ground_truth):
        rgb_gt = ground_truth['rgb'].to(model_outputs['rgb_values'].device)
        net_mask = model_outputs['network_object_mask']
        obj_mask = model_outputs['object_mask']
        rgb_loss = self.get_rgb_loss(model_outputs['rgb_values'], rgb_gt, net_mask, obj_mask)
        mask_loss = self.get_mask_loss(model_outputs['sdf_output'], net_mask, obj_mask)
        eikonal_loss = self.get_eikonal_loss(model_outputs['grad_theta'])
        loss = rgb_loss + self.eikonal_weight * eikonal_loss + self.mask_weight * mask_loss
        return {'loss': loss, 'rgb_loss': rgb_loss, 'eikonal_loss': eikonal_loss, 'mask_loss': mask_loss}
