"""Spectral-normalised video discriminator of the TAI training step (reference:
src/discriminators/SNDiscriminator.py).  Training-step plumbing; the only kernel it reaches is the bias +
LeakyReLU epilogue of its convolutions (models/layers.py:FusedSequential).

Behaviour kept from the reference: every forward runs ``Ip`` power iterations, divides ``weight.data`` by
the estimated top singular value IN PLACE (SNDiscriminator.py:63-68, 87-92) and keeps the vector ``u``
outside the ``state_dict``.  One change for data-parallel runs: ``u`` is drawn from a generator with a
fixed seed instead of the global RNG, so all ranks normalise with the same ``u`` and their replicas stay
bit-identical without a broadcast."""
from math import floor

import torch
import torch.nn as nn
from torch.nn import functional as F

from .. import ops
from ..models.layers import FusedSequential


def _l2normalize(v, eps=1e-12):
    if v.is_cuda and v.dtype == torch.float32 and not (torch.is_grad_enabled() and v.requires_grad):
        return ops.l2_normalize(v.contiguous(), eps)    # one launch instead of five
    return v / (((v ** 2).sum()) ** 0.5 + eps)


def max_singular_value(W, u=None, Ip=1):
    """Power iteration (SNDiscriminator.py:10-25).  Returns (sigma, u)."""
    if u is None:
        g = torch.Generator().manual_seed(W.size(0) * 7919 + W.size(1))
        u = torch.randn(1, W.size(0), generator=g).to(W.device)
    _u = u
    with torch.no_grad():
        for _ in range(Ip):
            _v = _l2normalize(torch.matmul(_u, W.data), eps=1e-12)
            _u = _l2normalize(torch.matmul(_v, W.data.t()), eps=1e-12)
        sigma = torch.matmul(torch.matmul(_v, W.data.t()), _u.t())
    return sigma, _u


class SNConv2d(nn.Conv2d):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 Ip=1):
        super(SNConv2d, self).__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias)
        self.Ip = Ip
        self.u = None

    def normalized_weight(self):
        """One spectral-norm update (power iteration + in-place division of ``weight.data``); returns the weight."""
        sigma, self.u = max_singular_value(self.weight.view(self.weight.size(0), -1), self.u, Ip=self.Ip)
        self.weight.data = self.weight.data / sigma
        return self.weight

    def forward(self, input):
        return F.conv2d(input, self.normalized_weight(), self.bias, self.stride, self.padding, self.dilation,
                        self.groups)


class SNLinear(nn.Linear):
    def __init__(self, in_features, out_features, bias=True, Ip=1):
        super(SNLinear, self).__init__(in_features, out_features, bias)
        self.u = None
        self.Ip = Ip

    def forward(self, input):
        sigma, self.u = max_singular_value(self.weight, self.u, Ip=self.Ip)
        self.weight.data = self.weight.data / sigma
        return F.linear(input, self.weight, self.bias)


class SNDiscriminator(nn.Module):
    """Slides a ``window_size``-frame window over the video: [B,T,C,H,W] -> [B, T-window_size+1] logits
    (SNDiscriminator.py:95-159)."""

    def __init__(self, img_size, c_dim, window_size, df_dim, Ip):
        super(SNDiscriminator, self).__init__()
        self.window_size = window_size
        h, w = img_size[0], img_size[1]
        layers = []
        cin = c_dim * window_size
        for mult in (1, 2, 4, 8):
            layers += [SNConv2d(cin, df_dim * mult, 4, stride=2, padding=1, Ip=Ip), nn.LeakyReLU(0.2)]
            cin = df_dim * mult
            h = floor((h + 2 * 1 - 4) / 2 + 1)
            w = floor((w + 2 * 1 - 4) / 2 + 1)
        self.conv_layers = FusedSequential(*layers)   # same children / keys; bias + LeakyReLU as one pass on CUDA
        self.num_sn_linear_in_feats = int(h * w * df_dim * 8)
        self.linear_layer = SNLinear(self.num_sn_linear_in_feats, 1, Ip=1)

    def forward(self, input):
        B, T, C, H, W = input.shape
        outs = []
        for t0 in range(T - self.window_size + 1):
            cur = input[:, t0:t0 + self.window_size].contiguous().view(B, self.window_size * C, H, W)
            feat = self.conv_layers(cur).view(B, self.num_sn_linear_in_feats)
            outs.append(self.linear_layer(feat))
        return torch.cat(outs, dim=1)
