"""``JointsMSELoss`` with the reference's registry name and signature (mmpose/models/losses/mse_loss.py:8-45).
On CUDA tensors the value (and the gradient w.r.t. the prediction) comes from the fused vpb_joints_mse_loss kernel
through a torch.autograd.Function; there is no CPU implementation."""
import torch
import torch.nn as nn

from .. import _lib
from .._lib import check, lib, ptr, stream_ptr
from ..builder import LOSSES


class _JointsMSEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, output, target, target_weight, loss_weight):
        if not output.is_cuda:
            raise _lib.VitposeLibError('JointsMSELoss runs on CUDA tensors only (no CPU fallback)')
        n, k = output.shape[:2]
        o = output.detach().reshape(n, k, -1).float().contiguous()
        t = target.detach().reshape(n, k, -1).float().contiguous()
        w = None if target_weight is None else target_weight.detach().reshape(n, k).float().contiguous()
        loss = torch.empty(1, device=o.device, dtype=torch.float32)
        grad = torch.empty_like(o) if output.requires_grad else None
        check(lib().vpb_joints_mse_loss(ptr(o), ptr(t), ptr(w), n, k, o.shape[2], float(loss_weight), ptr(loss),
                                        ptr(grad), stream_ptr()), 'vpb_joints_mse_loss')
        ctx.grad = grad
        ctx.shape = output.shape
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        return (ctx.grad.reshape(ctx.shape) * g if ctx.grad is not None else None), None, None, None


@LOSSES.register_module()
class JointsMSELoss(nn.Module):
    """MSE loss for heatmaps: mean over (N, H*W) per joint of ((pred - gt) * w)^2, summed over joints, / K."""

    def __init__(self, use_target_weight=False, loss_weight=1.):
        super().__init__()
        self.use_target_weight = use_target_weight
        self.loss_weight = loss_weight

    def forward(self, output, target, target_weight):
        w = target_weight if self.use_target_weight else None
        return _JointsMSEFn.apply(output, target, w, self.loss_weight)
