"""Parameter-free layers shared by the model families; each one stands for a library module of the
reference and routes it to a kernel of this library."""
import torch.nn as nn
import torch.nn.functional as F

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


class FusedSequential(nn.Sequential):
    """``nn.Sequential`` with the reference's children under the reference's names (state_dict keys unchanged) whose
    forward evaluates every ``nn.Conv2d`` / ``nn.ConvTranspose2d`` with bias [+ ``nn.ReLU`` / ``nn.LeakyReLU``] as a bias-free cuDNN
    convolution followed by ONE bias + activation pass of this library (``bias_act_forward_b200``, in place on
    the convolution output), instead of the library's broadcast ``add_``, ``clamp_min`` and, backward,
    ``threshold_backward`` and the bias-gradient ``sum``.  Same FP32 operations, same results.  CPU tensors take
    the plain route (the CPU port of the reference model)."""

    def forward(self, x):
        mods = list(self)
        i = 0
        while i < len(mods):
            m = mods[i]
            sn = getattr(m, 'normalized_weight', None)   # spectral-norm convolution: its weight update comes first
            if ((type(m) in (nn.Conv2d, nn.ConvTranspose2d) or (sn is not None and isinstance(m, nn.Conv2d)))
                    and m.bias is not None and x.is_cuda and m.padding_mode == 'zeros'
                    and x.dtype == m.weight.dtype):
                act, alpha, used = "none", 0.0, 1
                nxt = mods[i + 1] if i + 1 < len(mods) else None
                if type(nxt) is nn.ReLU:
                    act, used = "relu", 2
                elif type(nxt) is nn.LeakyReLU:
                    act, alpha, used = "leaky", nxt.negative_slope, 2
                if isinstance(m, nn.Conv2d):
                    w = sn() if sn is not None else m.weight
                    y = F.conv2d(x, w, None, m.stride, m.padding, m.dilation, m.groups)
                else:
                    y = F.conv_transpose2d(x, m.weight, None, m.stride, m.padding, m.output_padding, m.groups, m.dilation)
                x = ops.BiasActFunction.apply(y.contiguous(), m.bias, act, alpha)
                i += used
            else:
                x = m(x)
                i += 1
        return x
