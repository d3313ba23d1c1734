"""Losses of the TAI training step.

``GDL`` mirrors the reference module (src/losses/losses.py:4-45: same constructor, same result shapes).
``L2GDLLoss`` is what the training environments call: MSELoss + GDL of the inverse-transformed prediction
against the ground truth in ONE kernel pass each way (``l2_gdl_loss_forward/backward_b200``), replacing the
permute + inverse_transform + MSELoss + GDL chain of environments.py:363-371,447-451 (~25 launches forward
and as many backward per prediction tensor).
"""
import torch.nn as nn

from .. import ops


class GDL(nn.Module):
    """|dx(pred) - dx(target)| + |dy(pred) - dy(target)| on the common (H-1) x (W-1) support."""

    def __init__(self, reduce=True):
        super(GDL, self).__init__()
        self.reduce = reduce

    def forward(self, input, target):
        B = input.size(0)
        H, W = input.shape[-2:]
        lead = input.shape[:-2]
        x = input.reshape(-1, H, W)
        y = target.reshape(-1, H, W)
        # horizontal differences compared on rows 1.., vertical differences on columns 1.. (losses.py:30-35)
        w_term = ((x[:, :, :-1] - x[:, :, 1:]) - (y[:, :, :-1] - y[:, :, 1:])).abs()[:, 1:, :]
        h_term = ((x[:, 1:, :] - x[:, :-1, :]) - (y[:, 1:, :] - y[:, :-1, :])).abs()[:, :, 1:]
        loss = (w_term + h_term).reshape(*lead, H - 1, W - 1)
        return loss.reshape(B, -1).mean() if self.reduce else loss


class L2GDLLoss(nn.Module):
    """(mse, gdl) of ``inverse_transform(pred)`` vs ``inverse_transform(target)``; both are means, so the
    reference's time-major regrouping of the frames (environments.py:363-369) does not change them and is
    skipped.  CUDA tensors only -- there is no CPU version of the fused path."""

    def forward(self, pred, target):
        return ops.l2_gdl_loss(pred, target.detach(), 1.0, 0.5)   # util.py:22-23: (v + 1) / 2
