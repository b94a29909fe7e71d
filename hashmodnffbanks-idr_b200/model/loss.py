"""IDR loss with the reference's interface (model/loss.py:5-71): masked L1 on rgb, eikonal term and
the alpha-scaled mask BCE, all normalised by the number of rays.  Scalar reductions over [N] / [N,3]
tensors (negligible work, kept as device tensor ops so they sit on the autograd tape)."""
import torch
from torch import nn
from torch.nn import functional as F

from .. import kernels as K


class _FusedIDRLoss(torch.autograd.Function):
    """All three terms, their weighted sum and d loss / d (rgb_values, sdf_output, grad_theta) from ONE launch
    (csrc/render_glue.cu idr_loss_kernel; the eager formulation below is ~40 launches)."""

    @staticmethod
    def forward(ctx, rgb_values, sdf_output, grad_theta, rgb_gt, net_mask, obj_mask, eik_w, mask_w, alpha):
        need = any(ctx.needs_input_grad[:3])
        out, d_rgb, d_sdf, d_g = K.idr_loss(rgb_values.detach(), rgb_gt, net_mask, obj_mask, sdf_output.detach(),
                                            grad_theta.detach() if grad_theta is not None else None, eik_w, mask_w, alpha, need)
        ctx.sdf_shape = sdf_output.shape
        ctx.save_for_backward(d_rgb, d_sdf, d_g)
        loss, parts = out[0], out[1:]
        ctx.mark_non_differentiable(parts)
        return loss, parts

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_loss, _g_parts):
        d_rgb, d_sdf, d_g = ctx.saved_tensors
        if d_rgb is None:
            return (None,) * 9
        a, b, c = K.scale3(g_loss, d_rgb, d_sdf, d_g)
        return (a if ctx.needs_input_grad[0] else None, b.reshape(ctx.sdf_shape) if ctx.needs_input_grad[1] else None,
                c if (ctx.needs_input_grad[2] and c is not None) else None, None, None, None, None, None, None)


class IDRLoss(nn.Module):
    def __init__(self, eikonal_weight, mask_weight, alpha):
        super().__init__()
        self.eikonal_weight = eikonal_weight
        self.mask_weight = mask_weight
        self.alpha = alpha
        self.fused = True           # one-launch loss + gradient; False = the eager tensor-op formulation below

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
        rgb, sdf, gth = model_outputs['rgb_values'], model_outputs['sdf_output'], model_outputs['grad_theta']
        if (self.fused and rgb.is_cuda and rgb.dtype == torch.float32 and gth is not None and gth.dim() == 2
                and gth.shape[0] > 0 and net_mask.dtype == torch.bool and obj_mask.dtype == torch.bool):
            loss, parts = _FusedIDRLoss.apply(rgb.reshape(-1, 3), sdf, gth, rgb_gt, net_mask, obj_mask,
                                              float(self.eikonal_weight), float(self.mask_weight), float(self.alpha))
            return {'loss': loss, 'rgb_loss': parts[0], 'eikonal_loss': parts[1], 'mask_loss': parts[2]}
        rgb_loss = self.get_rgb_loss(model_outputs['rgb_values'], rgb_gt, net_mask, obj_mask)
        mask_loss = self.get_mask_loss(model_outputs['sdf_output'], net_mask, obj_mask)
        eikonal_loss = self.get_eikonal_loss(model_outputs['grad_theta'])
        loss = rgb_loss + self.eikonal_weight * eikonal_loss + self.mask_weight * mask_loss
        return {'loss': loss, 'rgb_loss': rgb_loss, 'eikonal_loss': eikonal_loss, 'mask_loss': mask_loss}
