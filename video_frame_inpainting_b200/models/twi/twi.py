"""bi-TWI ablation: the TAI kernel network without the time-ratio input, followed by a time-weighted
blend.  Mirror of the reference's ``src/models/twi/twi.py``; it exists here because it is the same fused
kernel with different blend weights: pred_t = (1 - w_t) * Dot1 + w_t * Dot2 (twi.py:105),
w = linspace(0, 1, T+2)[1:-1] (twi.py:90)."""
import numpy as np
import torch.nn as nn
from torch.nn import functional as F

from ..mcnet.mcnet import MCNet, Residual
from ..tai.tai import TAI, TAIFillInModel


class TWI(TAI):
    """Time-agnostic kernel network (twi.py:123-231): ``create_decoder_blocks(..., rc_loc=-1)`` (twi.py:162)."""
    RC_LOC = -1

    def forward(self, variableInput1, variableInput2, variableDyn1, variableDyn2, variableCont1, variableCont2,
                variableRes):
        return super(TWI, self).forward(variableInput1, variableInput2, variableDyn1, variableDyn2, variableCont1,
                                        variableCont2, variableRes, ratio=0)


class TimeWeightedInterpolationFillInModel(TAIFillInModel):
    """Same forward pipeline as bi-TAI with attribute names ``mcnet`` / ``interp_net`` (twi.py:43-49)."""

    def __init__(self, gf_dim, c_dim, feature_size, ks, num_block=5, kf_dim=32, layers=3, forget_bias=1,
                 activation=F.tanh, bias=True):
        nn.Module.__init__(self)
        self.c_dim = c_dim
        self.conv_lstm_state_size = 8 * gf_dim
        self.mcnet = MCNet(gf_dim, c_dim, feature_size, forget_bias=forget_bias, activation=activation, bias=bias)
        self.merge_residual3 = Residual(gf_dim * 8, kf_dim * 4)
        self.merge_residual2 = Residual(gf_dim * 4, kf_dim * 2)
        self.merge_residual1 = Residual(gf_dim * 2, kf_dim * 1)
        self.interp_net = TWI(gf_dim, ks, num_block, layers, kf_dim)

    # the shared forward of TAIFillInModel addresses the sub-networks by the bi-TAI names
    @property
    def generator(self):
        return self.mcnet

    @property
    def kernelnet(self):
        return self.interp_net

    def blend_weights(self, T):
        w = np.linspace(0, 1, num=T + 2).tolist()[1:-1]
        return [(1 - w[t], w[t], 0) for t in range(T)]
