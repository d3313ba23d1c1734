"""Image gradient-difference loss used by the TAI training step (reference: src/losses/losses.py:4-45)."""
import torch.nn as nn


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
