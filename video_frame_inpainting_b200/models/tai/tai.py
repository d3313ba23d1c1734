"""bi-TAI: bidirectional MC-Net + TAI kernel network, with the filter-and-blend tail on sm_100a kernels.

Host-side mirror of the reference's ``src/models/tai/tai.py``: same classes, constructor signatures,
attribute names (``generator``, ``merge_residual{1,2,3}``, ``kernelnet``, ``moduleConv`` ...) and output
dict.  What changes on the hot path:

* reference, per middle frame: 2 x ReplicationPad2d -> 2 x SeparableConvolution -> 0.5*Dot1 + 0.5*Dot2
  (tai.py:229-236, 105): 2 pad kernels, 2 naive sepconv kernels, 3 elementwise kernels;
* here: ONE launch of ``tai_fused_forward_b200`` (pad folded into the halo load, both streams filtered,
  blend in the epilogue, Dot1/Dot2 still emitted because they are part of the output dict, tai.py:117-118).
  ``TAI.forward`` keeps its reference contract (returns Dot1, Dot2) and ``TAI.filter_and_blend`` is the
  fused entry that ``TAIFillInModel.forward`` uses.

torch-0.3.1 semantics pinned here: ``nn.Upsample(mode='bilinear')`` interpolated with the
align-corners mapping, so every bilinear upsample is ``align_corners=True``; ``xrange`` / py2 ``/`` are
``range`` / ``//``.
"""
import math

import numpy as np
import torch
import torch.nn as nn
from torch.nn import functional as F

from ... import ops
from ..layers import BilinearUp2, FusedSequential
from ...separable_convolution.SeparableConvolution import SeparableConvolution
from ..mcnet.mcnet import MCNet, Residual, gray_difference_frames


def _up2():
    return BilinearUp2()  # torch-0.3.1 bilinear mapping (align corners), one kernel of this library


class TAIFillInModel(nn.Module):
    """``forward(T, preceding_frames[B,K,C,H,W], following_frames[B,F,C,H,W])`` -> dict with keys
    pred, pred_forward, pred_backward, interp_net_outputs_1, interp_net_outputs_2   (tai.py:14-120)."""

    def __init__(self, gf_dim, c_dim, feature_size, ks, num_block=5, kf_dim=32, layers=3, forget_bias=1,
                 activation=F.tanh, bias=True):
        super(TAIFillInModel, self).__init__()
        self.c_dim = c_dim
        self.conv_lstm_state_size = 8 * gf_dim
        self.generator = MCNet(gf_dim, c_dim, feature_size, forget_bias=forget_bias, activation=activation, bias=bias)
        self.merge_residual3 = Residual(gf_dim * 8, kf_dim * 4)
        self.merge_residual2 = Residual(gf_dim * 4, kf_dim * 2)
        self.merge_residual1 = Residual(gf_dim * 2, kf_dim * 1)
        self.kernelnet = TAI(gf_dim, ks, num_block, layers, kf_dim)

    # run the forward and the backward MC-Net stream as one pass over 2B clips when K == F, and the kernel
    # network once over the T*B middle frames instead of once per t (see forward()); the CPU port of the
    # reference switches both off
    batch_streams = True
    batch_time = True

    def blend_weights(self, T):
        """(a_t, b_t) of pred_t = a_t*Dot1 + b_t*Dot2 and the time ratio fed to the kernel net.
        bi-TAI: a = b = 0.5 (tai.py:105), ratio_t = 1 - w_t, w = linspace(0,1,T+2)[1:-1] (tai.py:90,99)."""
        w = np.linspace(0, 1, num=T + 2).tolist()[1:-1]
        return [(0.5, 0.5, 1 - w[t]) for t in range(T)]

    def forward(self, T, preceding_frames, following_frames):
        K = preceding_frames.size(1)
        F_ = following_frames.size(1)
        xt = preceding_frames[:, -1]
        xt_F = following_frames[:, 0]
        diff_in = gray_difference_frames(preceding_frames)
        diff_in_F = gray_difference_frames(following_frames, reverse=True)  # time-reversed (tai.py:71-74)

        weights = self.blend_weights(T)
        B = xt.size(0)
        if self.batch_streams and K == F_:
            # The two MC-Net passes share their weights and never mix samples, so they are ONE pass over the
            # 2B clips [preceding; time-reversed following] (the reference runs them back to back,
            # tai.py:77-84): half the launches, and at small batches twice the tiles per kernel (batch-1
            # inference ran 39 ms of kernels of <= 64 tiles on 148 SMs).
            both = self.generator(K, T, torch.cat([diff_in, diff_in_F], 0), torch.cat([xt, xt_F], 0))
            if self.batch_time and T > 1 and len(set((w[0], w[1]) for w in weights)) == 1 and \
                    ops.gather_concat_ok(*both[0], *both[1], *both[2], *[r for res_t in both[3] for r in res_t]):
                return self._forward_batched(T, B, both, weights)

            # unbind() hands each stream its half as a contiguous view; its backward is a single stack
            def halves(x):
                return x.view(2, B, *x.shape[1:]).unbind(0)

            forward_pred, backward_pred = zip(*[halves(x) for x in both[0]])
            forward_dyn, backward_dyn = zip(*[halves(x) for x in both[1]])
            forward_cont, backward_cont = zip(*[halves(x) for x in both[2]])
            res_pairs = [[halves(r) for r in res_t] for res_t in both[3]]
            forward_res = [[r[0] for r in res_t] for res_t in res_pairs]
            backward_res = [[r[1] for r in res_t] for res_t in res_pairs]
            forward_pred, forward_dyn, forward_cont = list(forward_pred), list(forward_dyn), list(forward_cont)
            backward_pred, backward_dyn, backward_cont = list(backward_pred), list(backward_dyn), list(backward_cont)
        else:
            forward_pred, forward_dyn, forward_cont, forward_res = self.generator(K, T, diff_in, xt)
            backward_pred, backward_dyn, backward_cont, backward_res = self.generator(F_, T, diff_in_F, xt_F)
        backward_pred, backward_dyn = backward_pred[::-1], backward_dyn[::-1]
        backward_cont, backward_res = backward_cont[::-1], backward_res[::-1]

        combination, outputs_1, outputs_2 = [], [], []
        if self.batch_time and T > 1:
            # The T kernel-net evaluations depend on MC-Net outputs only, not on each other (tai.py:91-105 runs
            # them in a loop): ONE pass over the T*B samples [t = 0; t = 1; ...], the time ratio as a per-block
            # constant plane, then the fused pad + sepconv + blend kernel over all T*B frames (per t when the
            # blend weights differ, bi-TWI).

            def cat(xs):
                return torch.cat(list(xs), 0)

            merged_res = [self.merge_residual1(cat(r[0] for r in forward_res), cat(r[0] for r in backward_res)),
                          self.merge_residual2(cat(r[1] for r in forward_res), cat(r[1] for r in backward_res)),
                          self.merge_residual3(cat(r[2] for r in forward_res), cat(r[2] for r in backward_res))]
            v1, h1, v2, h2 = self.kernelnet.kernel_maps(cat(forward_dyn), cat(backward_dyn), cat(forward_cont),
                                                        cat(backward_cont), merged_res, ratio=[w[2] for w in weights])
            pf, pb = cat(forward_pred), cat(backward_pred)
            if len(set((w[0], w[1]) for w in weights)) == 1:
                pred, dot1, dot2 = self.kernelnet.apply_maps(pf, pb, v1, h1, v2, h2, weights[0][0], weights[0][1])
                combination = list(pred.view(T, B, *pred.shape[1:]).unbind(0))
                outputs_1 = list(dot1.view(T, B, *dot1.shape[1:]).unbind(0))
                outputs_2 = list(dot2.view(T, B, *dot2.shape[1:]).unbind(0))
            else:
                for t, (a, b, _) in enumerate(weights):
                    sl = slice(t * B, (t + 1) * B)
                    pred_t, dot1, dot2 = self.kernelnet.apply_maps(pf[sl], pb[sl], v1[sl], h1[sl], v2[sl], h2[sl], a, b)
                    combination.append(pred_t)
                    outputs_1.append(dot1)
                    outputs_2.append(dot2)
        else:
            for t, (a, b, ratio) in enumerate(weights):
                merged_res = [self.merge_residual1(forward_res[t][0], backward_res[t][0]),
                              self.merge_residual2(forward_res[t][1], backward_res[t][1]),
                              self.merge_residual3(forward_res[t][2], backward_res[t][2])]
                pred_t, dot1, dot2 = self.kernelnet.filter_and_blend(
                    forward_pred[t], backward_pred[t], forward_dyn[t], backward_dyn[t], forward_cont[t],
                    backward_cont[t], merged_res, ratio=ratio, a=a, b=b)
                combination.append(pred_t)
                outputs_1.append(dot1)
                outputs_2.append(dot2)

        return {
            'pred': torch.stack(combination, dim=1),
            'pred_forward': torch.stack(forward_pred, dim=1),
            'pred_backward': torch.stack(backward_pred, dim=1),
            'interp_net_outputs_1': torch.stack(outputs_1, dim=1),
            'interp_net_outputs_2': torch.stack(outputs_2, dim=1),
        }


    def _forward_batched(self, T, B, both, weights):
        """Both MC-Net streams as one 2B batch AND the kernel network once over the T*B middle frames: every input
        of the kernel network is gathered straight from the per-step tensors of the 2B pass -- sample (t, b) takes
        the forward stream of step t (samples 0..B-1) and the backward stream of step T-1-t (samples B..2B-1), the
        time reversal of tai.py:85-88 -- in ONE launch per tensor (``gather_concat_*``), where stacking the T steps
        of each stream and then concatenating the streams along channels would copy everything twice."""
        pred_l, dyn_l, cont_l, res_l = both

        def merge(xs):   # T x [2B,C,h,w] -> [T*B, 2C, h, w]: (forward stream at t | backward stream at T-1-t)
            C = xs[0].shape[1]
            blocks = []
            for t in range(T):
                blocks.append((t, 0, B, t * B, 1, 0, C, 0.0))
                blocks.append((T - 1 - t, B, B, t * B, 1, C, C, 0.0))
            return ops.GatherConcatFunction.apply((T * B, 2 * C, tuple(blocks)), *xs)

        merged_res = [self.merge_residual1.res(merge([r[0] for r in res_l])),
                      self.merge_residual2.res(merge([r[1] for r in res_l])),
                      self.merge_residual3.res(merge([r[2] for r in res_l]))]
        # kernel-network input: cat(dyn1, dyn2, cont1, cont2) of tai.py:188
        Cd, Cc = dyn_l[0].shape[1], cont_l[0].shape[1]
        blocks = []
        for t in range(T):
            blocks += [(t, 0, B, t * B, 1, 0, Cd, 0.0), (T - 1 - t, B, B, t * B, 1, Cd, Cd, 0.0),
                       (T + t, 0, B, t * B, 1, 2 * Cd, Cc, 0.0), (2 * T - 1 - t, B, B, t * B, 1, 2 * Cd + Cc, Cc, 0.0)]
        x0 = ops.GatherConcatFunction.apply((T * B, 2 * Cd + 2 * Cc, tuple(blocks)), *dyn_l, *cont_l)
        v1, h1, v2, h2 = self.kernelnet.kernel_maps_from(x0, merged_res, ratio=[w[2] for w in weights])
        # the two predictions per middle frame, [pf; pb] as one tensor of 2*T*B samples
        C = pred_l[0].shape[1]
        blocks = []
        for t in range(T):
            blocks += [(t, 0, B, t * B, 1, 0, C, 0.0), (T - 1 - t, B, B, (T + t) * B, 1, 0, C, 0.0)]
        pfb = ops.GatherConcatFunction.apply((2 * T * B, C, tuple(blocks)), *pred_l)
        pred, dot1, dot2 = self.kernelnet.apply_maps(pfb[:T * B], pfb[T * B:], v1, h1, v2, h2, weights[0][0], weights[0][1])
        # outputs [B,T,C,H,W]: pred_forward[b,t] = stream 0 of step t, pred_backward[b,t] = stream 1 of step T-1-t
        blocks = []
        for t in range(T):
            blocks += [(t, 0, B, t, T, 0, C, 0.0), (T - 1 - t, B, B, B * T + t, T, 0, C, 0.0)]
        fb = ops.GatherConcatFunction.apply((2 * B * T, C, tuple(blocks)), *pred_l)
        HW = pred.shape[2:]

        def bt(x):    # [T*B,C,H,W] (t-major) -> [B,T,C,H,W]
            return x.view(T, B, C, *HW).transpose(0, 1).contiguous()
        return {
            'pred': bt(pred),
            'pred_forward': fb[:B * T].view(B, T, C, *HW),
            'pred_backward': fb[B * T:].view(B, T, C, *HW),
            'interp_net_outputs_1': bt(dot1),
            'interp_net_outputs_2': bt(dot2),
        }


class TAI(nn.Module):
    """Kernel network: encoder / decoder over [dyn1, dyn2, cont1, cont2] -> four 1-D kernel maps
    V1, H1, V2, H2 [B, ks, H, W], applied to the two predictions   (tai.py:123-237)."""

    RC_LOC = 4  # decoder block (1-based) that receives the time-ratio plane; TWI overrides it with -1

    def __init__(self, gf_dim, ks, num_block, layers, kf_dim):
        super(TAI, self).__init__()
        assert layers >= 1, 'layers in per block should be no smaller than 1, but layers=[%d]' % layers
        assert num_block >= 4, '# blocks should be no less than 3, but num_block=%d' % num_block
        self.kf_dim = kf_dim
        self.ks = ks
        self.layers = layers
        self.num_block = num_block
        self.rc_loc = self.RC_LOC

        moduleConv, modulePool = create_encoder_blocks(3, num_block, layers, gf_dim * 8 * 2, kf_dim)
        self.moduleConv = nn.ModuleList(moduleConv)
        self.modulePool = nn.ModuleList(modulePool)
        moduleDeconv, moduleUpsample = create_decoder_blocks(num_block - 1, kf_dim, layers, self.rc_loc)
        self.moduleDeconv = nn.ModuleList(moduleDeconv)
        self.moduleUpsample = nn.ModuleList(moduleUpsample)

        self.moduleVertical1 = create_1d_kernel_generator_block(layers, kf_dim, ks)
        self.moduleVertical2 = create_1d_kernel_generator_block(layers, kf_dim, ks)
        self.moduleHorizontal1 = create_1d_kernel_generator_block(layers, kf_dim, ks)
        self.moduleHorizontal2 = create_1d_kernel_generator_block(layers, kf_dim, ks)

        pad = int(math.floor(ks / 2.0))
        self.modulePad = nn.ReplicationPad2d([pad, pad, pad, pad])
        self.separableConvolution = SeparableConvolution.apply

    def kernel_maps(self, variableDyn1, variableDyn2, variableCont1, variableCont2, variableRes, ratio=0):
        """Everything of tai.py:188-226,230-235 that produces V1, H1, V2, H2."""
        x = torch.cat([variableDyn1, variableDyn2, variableCont1, variableCont2], 1)
        return self.kernel_maps_from(x, variableRes, ratio)

    def kernel_maps_from(self, x, variableRes, ratio=0):
        """kernel_maps for an already assembled input cat(dyn1, dyn2, cont1, cont2)."""
        nb = self.num_block
        enc = []
        for i in range(nb - 3):
            enc.append(self.moduleConv[i](x))
            x = self.modulePool[i](enc[-1])
        for i in range(nb - 1):
            x = self.moduleDeconv[i](x)
            if i == self.rc_loc - 1:  # time-ratio plane; reachable only when num_block >= 5 (tai.py:213-217)
                ratios = list(ratio) if isinstance(ratio, (list, tuple)) else [ratio]
                n = x.size(0) // len(ratios)      # batch = len(ratios) equal blocks, one ratio per block
                if x.is_cuda and (x.size(2) * x.size(3)) % 4 == 0:
                    # x and the constant plane(s) in one launch (fill blocks), one launch for the adjoint
                    Cx = x.size(1)
                    blocks = [(0, 0, x.size(0), 0, 1, 0, Cx, 0.0)]
                    blocks += [(None, 0, n, i * n, 1, Cx, 1, float(r)) for i, r in enumerate(ratios)]
                    x = ops.GatherConcatFunction.apply((x.size(0), Cx + 1, tuple(blocks)), x.contiguous())
                else:
                    plane = torch.cat([x.new_full((n, 1, x.size(2), x.size(3)), float(r)) for r in ratios], 0)
                    x = torch.cat([x, plane], dim=1)
            x = self.moduleUpsample[i](x)
            x = x + (enc[nb - 3 - i - 1] if i < nb - 3 else variableRes[nb - i - 1])
        return (self.moduleVertical1(x), self.moduleHorizontal1(x), self.moduleVertical2(x), self.moduleHorizontal2(x))

    def forward(self, variableInput1, variableInput2, variableDyn1, variableDyn2, variableCont1, variableCont2,
                variableRes, ratio=0):
        """Reference contract: returns (Dot1, Dot2), the two filtered predictions (tai.py:174-237), through
        the reference-shaped operator ``SeparableConvolution.apply(pad(x), V, H, ks)``."""
        v1, h1, v2, h2 = self.kernel_maps(variableDyn1, variableDyn2, variableCont1, variableCont2, variableRes, ratio)
        dot1 = self.separableConvolution(self.modulePad(variableInput1).contiguous(), v1, h1, self.ks)
        dot2 = self.separableConvolution(self.modulePad(variableInput2).contiguous(), v2, h2, self.ks)
        return dot1, dot2

    def filter_and_blend(self, variableInput1, variableInput2, variableDyn1, variableDyn2, variableCont1,
                         variableCont2, variableRes, ratio=0, a=0.5, b=0.5):
        """Fused tail: (a*Dot1 + b*Dot2, Dot1, Dot2) in one kernel launch (pad + 2 x sepconv + blend)."""
        v1, h1, v2, h2 = self.kernel_maps(variableDyn1, variableDyn2, variableCont1, variableCont2, variableRes, ratio)
        return self.apply_maps(variableInput1, variableInput2, v1, h1, v2, h2, a, b)

    def apply_maps(self, variableInput1, variableInput2, v1, h1, v2, h2, a=0.5, b=0.5):
        """(a*Dot1 + b*Dot2, Dot1, Dot2) for given kernel maps: replication pad, the two separable convolutions and
        the blend as one kernel launch (tai.py:229-236 + 105)."""
        return ops.tai_blend_sepconv(variableInput1.contiguous(), variableInput2.contiguous(), v1.contiguous(),
                                     h1.contiguous(), v2.contiguous(), h2.contiguous(), self.ks, a, b)


# ------------------------------------------------------------------------------------------------
# builders (tai.py:244-347); module indices inside each Sequential match the reference's
# ------------------------------------------------------------------------------------------------

def create_basic_conv_block(num_layers, num_in_channels, num_out_channels):
    """num_layers x (3x3 conv, ReLU), resolution preserving   (tai.py:244-263)."""
    seq = []
    cin = num_in_channels
    for _ in range(num_layers):
        seq += [nn.Conv2d(cin, num_out_channels, kernel_size=3, stride=1, padding=1), nn.ReLU(inplace=False)]
        cin = num_out_channels
    return FusedSequential(*seq)


def create_1d_kernel_generator_block(num_layers, kf_dim, ks):
    """(num_layers-1) x conv(2kf->2kf)+ReLU, conv(2kf->ks)+ReLU, bilinear x2, conv(ks->ks); the output has
    no normalisation and may be negative   (tai.py:266-286)."""
    seq = []
    for i in range(num_layers):
        cout = ks if i == num_layers - 1 else kf_dim * 2
        seq += [nn.Conv2d(kf_dim * 2, cout, kernel_size=3, stride=1, padding=1), nn.ReLU(inplace=False)]
    seq += [_up2(), nn.Conv2d(ks, ks, kernel_size=3, stride=1, padding=1)]
    return FusedSequential(*seq)


def create_encoder_blocks(start_i, end_i, layers, if_dim, kf_dim):
    """Blocks i = start_i..end_i-1 with kf_dim * 2**i channels, each followed by AvgPool2d(2) (tai.py:289-310)."""
    convs, pools = [], []
    cin = if_dim
    for i in range(start_i, end_i):
        cout = kf_dim * (2 ** i)
        convs.append(create_basic_conv_block(layers, cin, cout))
        pools.append(nn.AvgPool2d(kernel_size=2, stride=2))
        cin = cout
    return convs, pools


def create_decoder_blocks(num_block, kf_dim, layers, rc_loc):
    """Decoder block i: conv block, then bilinear x2 + conv + ReLU; block rc_loc-1 takes one extra input
    channel (the time ratio)   (tai.py:313-347)."""
    deconvs, upsamples = [], []
    for i in range(num_block):
        c_out = kf_dim * 2 ** (num_block - i)
        c_in = c_out if i == 0 else kf_dim * 2 ** (num_block - i + 1)
        deconvs.append(create_basic_conv_block(layers, c_in, c_out))
        extra = 1 if i == rc_loc - 1 else 0
        upsamples.append(FusedSequential(_up2(),
                                       nn.Conv2d(c_out + extra, c_out, kernel_size=3, stride=1, padding=1),
                                       nn.ReLU(inplace=False)))
    return deconvs, upsamples
