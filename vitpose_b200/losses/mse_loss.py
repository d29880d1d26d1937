"""``JointsMSELoss`` registry entry (mmpose/models/losses/mse_loss.py:8-45) so that ``loss_keypoint`` blocks of
ViTPose configs build unchanged.  Training-step row of SURVEY.md §8 (a17): the forward value is computed with
torch ops on the tensors' device; a fused CUDA fwd/bwd kernel belongs to the training-step milestone
(DESIGN.md "what comes next")."""
import torch.nn as nn

from ..builder import LOSSES


@LOSSES.register_module()
class JointsMSELoss(nn.Module):
    def __init__(self, use_target_weight=False, loss_weight=1.):
        super().__init__()
        self.use_target_weight = use_target_weight
        self.loss_weight = loss_weight

    def forward(self, output, target, target_weight):
        n, k = output.size(0), output.size(1)
        pred = output.reshape(n, k, -1)
        gt = target.reshape(n, k, -1)
        if self.use_target_weight:
            pred = pred * target_weight
            gt = gt * target_weight
        # sum over joints of the per-joint mean squared error, / K  (mse_loss.py:35-45)
        return ((pred - gt) ** 2).mean(dim=(0, 2)).sum() / k * self.loss_weight
