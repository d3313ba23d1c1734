"""Parameter-free layers shared by the model families; each one stands for a library module of the
reference and routes it to a kernel of this library."""
import torch.nn as nn

from .. import ops


class BilinearUp2(nn.Module):
    """``nn.Upsample(scale_factor=2, mode='bilinear')`` as the reference's torch 0.3.1 evaluated it (the
    align-corners mapping; tai.py:283,337,343; slomo.py:113-149).  Holds no parameters or buffers, so the
    state_dict keys and the module indices inside each ``nn.Sequential`` are those of the reference."""

    def forward(self, x):
        return ops.UpsampleBilinear2xFunction.apply(x.contiguous())

    def extra_repr(self):
        return "scale_factor=2, mode=bilinear (align-corners mapping of torch 0.3.1)"
