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


class MaxPool2(nn.Module):
    """``nn.MaxPool2d(2)`` (mcnet.py:28-45; slomo.py:47-85) through the library's 2 x 2 pooling kernels: the
    selected position is kept as a one-byte code instead of an int64 index and the backward kernel writes the
    input gradient in one pass (the library's pair was 2.3 % of the KTH training step).  Same tie rule as the
    library kernel (first maximum in scan order).  Parameter-free: state_dict keys and Sequential indices are
    those of the reference."""

    def forward(self, x):
        return ops.MaxPool2x2Function.apply(x)

    def extra_repr(self):
        return "kernel_size=2, stride=2"
