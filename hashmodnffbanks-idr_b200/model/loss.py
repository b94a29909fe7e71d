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

    def get_rgb_loss(self, rgb_values, rgb_gt, network_object_mask, object_mask):
        both = network_object_mask & object_mask
        sel = rgb_values[both]
        if sel.shape[0] == 0:
            return torch.zeros((), device=rgb_values.device)
        gt = rgb_gt.reshape(-1, 3)[both]
        return (sel - gt).abs().sum() / float(object_mask.shape[0])

    def get_eikonal_loss(self, grad_theta):
        if grad_theta.shape[0] == 0:
            return torch.zeros((), device=grad_theta.device)
        return ((grad_theta.norm(2, dim=1) - 1) ** 2).mean()

    def get_mask_loss(self, sdf_output, network_object_mask, object_mask):
        neg = ~(network_object_mask & object_mask)
        logits = -self.alpha * sdf_output[neg]
        if logits.shape[0] == 0:
            return torch.zeros((), device=sdf_output.device)
        gt = object_mask[neg].float()
        bce = F.binary_cross_entropy_with_logits(logits.squeeze(-1), gt, reduction='sum')
        return (1 / self.alpha) * bce / float(object_mask.shape[0])

    def forward(self, model_outputs, ground_truth):
        rgb_gt = ground_truth['rgb'].to(model_outputs['rgb_values'].device)
        net_mask = model_outputs['network_object_mask']
        obj_mask = model_outputs['object_mask']
        rgb_loss = self.get_rgb_loss(model_outputs['rgb_values'], rgb_gt, net_mask, obj_mask)
        mask_loss = self.get_mask_loss(model_outputs['sdf_output'], net_mask, obj_mask)
        eikonal_loss = self.get_eikonal_loss(model_outputs['grad_theta'])
        loss = rgb_loss + self.eikonal_weight * eikonal_loss + self.mask_weight * mask_loss
        return {'loss': loss, 'rgb_loss': rgb_loss, 'eikonal_loss': eikonal_loss, 'mask_loss': mask_loss}
