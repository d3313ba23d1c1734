"""MC-Net video predictor with the ConvLSTM gate path on the sm_100a kernel.

Host-side mirror of the reference's ``src/models/mcnet/mcnet.py`` (torch 0.3.1 / Python 2) for the
TAI hot path: same class names, constructor signatures, attribute names (so reference ``state_dict``s
load) and forward contracts.  The convolutions stay on cuDNN; what changes is the recurrent glue:

* ``ConvLstmCell.forward`` (mcnet.py:281-294): the chunk / sigmoid / tanh / cat chain (~10 elementwise
  launches, 100+ B moved per state element) is one kernel, ``convlstm_gates_forward_b200`` (28 B per
  element), with its own backward;
* Python-2 integer divisions of the reference (mcnet.py:278, 384) are written ``//``.
"""
import torch
import torch.nn as nn
from torch.nn import functional as F

from ... import ops
from ..layers import FusedSequential, MaxPool2
from ...util.util import bgr2gray, bgr2gray_batched, inverse_transform


def _conv_relu_chain(channels, kernel, transposed=False, last=None):
    """[Conv(c0->c1,k), ReLU, Conv(c1->c2,k), ReLU, ...]; `last` replaces the final activation."""
    conv = nn.ConvTranspose2d if transposed else nn.Conv2d
    layers = []
    for idx, (cin, cout) in enumerate(zip(channels[:-1], channels[1:])):
        layers.append(conv(cin, cout, kernel, padding=kernel // 2))
        is_last = idx == len(channels) - 2
        layers.append(last if (is_last and last is not None) else nn.ReLU())
    return layers


class MotionEnc(nn.Module):
    """Difference frame [B,1,H,W] -> ([B,4*gf,H/8,W/8], three skip activations)   (mcnet.py:14-60)."""

    def __init__(self, gf_dim):
        super(MotionEnc, self).__init__()
        self.dyn_conv1 = FusedSequential(nn.Conv2d(1, gf_dim, 5, padding=2), nn.ReLU())
        self.dyn_conv2 = FusedSequential(MaxPool2(), nn.Conv2d(gf_dim, gf_dim * 2, 5, padding=2), nn.ReLU())
        self.dyn_conv3 = FusedSequential(MaxPool2(), nn.Conv2d(gf_dim * 2, gf_dim * 4, 7, padding=3), nn.ReLU())
        self.pool3 = MaxPool2()

    def forward(self, input_diff):
        skips = []
        x = input_diff
        for stage in (self.dyn_conv1, self.dyn_conv2, self.dyn_conv3):
            x = stage(x)
            skips.append(x)
        return self.pool3(x), skips


class ContentEnc(nn.Module):
    """Frame [B,c,H,W] -> ([B,4*gf,H/8,W/8], three skip activations)   (mcnet.py:63-119)."""

    def __init__(self, c_dim, gf_dim):
        super(ContentEnc, self).__init__()
        self.cont_conv1 = FusedSequential(*_conv_relu_chain([c_dim, gf_dim, gf_dim], 3))
        self.cont_conv2 = FusedSequential(MaxPool2(), *_conv_relu_chain([gf_dim, gf_dim * 2, gf_dim * 2], 3))
        self.cont_conv3 = FusedSequential(MaxPool2(),
                                        *_conv_relu_chain([gf_dim * 2, gf_dim * 4, gf_dim * 4, gf_dim * 4], 3))
        self.pool3 = MaxPool2()

    def forward(self, raw):
        skips = []
        x = raw
        for stage in (self.cont_conv1, self.cont_conv2, self.cont_conv3):
            x = stage(x)
            skips.append(x)
        return self.pool3(x), skips


def _cat_channels(a, b):
    """torch.cat((a, b), dim=1) -- through the library's gather kernel on the GPU (one launch each way; the adjoint
    writes two contiguous gradients instead of returning strided slices)."""
    if ops.gather_concat_ok(a, b):
        return ops.cat_channels((a, b))
    return torch.cat((a, b), dim=1)


class CombLayers(nn.Module):
    """cat(h_dyn, h_cont) -> three 3x3 convolutions   (mcnet.py:122-153)."""

    def __init__(self, gf_dim):
        super(CombLayers, self).__init__()
        self.h_comb = FusedSequential(*_conv_relu_chain([gf_dim * 8, gf_dim * 4, gf_dim * 2, gf_dim * 4], 3))

    def forward(self, h_dyn, h_cont):
        return self.h_comb(_cat_channels(h_dyn, h_cont))


class Residual(nn.Module):
    """cat(dyn, cont) -> conv, ReLU, conv   (mcnet.py:156-185)."""

    def __init__(self, in_dim, out_dim):
        super(Residual, self).__init__()
        self.res = FusedSequential(nn.Conv2d(in_dim, out_dim, 3, padding=1), nn.ReLU(),
                                 nn.Conv2d(out_dim, out_dim, 3, padding=1))

    def forward(self, input_dyn, input_cont):
        return self.res(_cat_channels(input_dyn, input_cont))


class DecCnn(nn.Module):
    """Decoder: three (zero-insertion unpool + residual add + transposed-conv block) stages, tanh output
    (mcnet.py:188-256)."""

    def __init__(self, c_dim, gf_dim):
        super(DecCnn, self).__init__()
        self.dec3 = FusedSequential(*_conv_relu_chain([gf_dim * 4, gf_dim * 4, gf_dim * 4, gf_dim * 2], 3, transposed=True))
        self.dec2 = FusedSequential(*_conv_relu_chain([gf_dim * 2, gf_dim * 2, gf_dim], 3, transposed=True))
        self.dec1 = FusedSequential(*_conv_relu_chain([gf_dim, gf_dim, c_dim], 3, transposed=True, last=nn.Tanh()))

    def forward(self, comb, res1, res2, res3):
        dec3_out = self.dec3(self.unpool_add(comb, res3))
        dec2_out = self.dec2(self.unpool_add(dec3_out, res2))
        return self.dec1(self.unpool_add(dec2_out, res1))

    def unpool_add(self, x, res):
        """``fixed_unpooling(x) + res`` (mcnet.py:234-236) as one kernel."""
        return ops.UnpoolAddFunction.apply(x.contiguous(), res.contiguous())

    def fixed_unpooling(self, x):
        """out[2y, 2x] = x[y, x], the other three elements of every 2x2 cell are zero (mcnet.py:240-256;
        the reference builds it from two cats, a clone().zero_() and two permutes)."""
        B, C, H, W = x.shape
        out = x.new_zeros(B, C, 2 * H, 2 * W)
        out[:, :, ::2, ::2] = x
        return out


def _is_tanh(fn):
    return fn in (torch.tanh, F.tanh) or getattr(fn, "__name__", "") == "tanh"


class ConvLstmCell(nn.Module):
    """Convolutional LSTM cell; ``forward(input, state) -> (new_h, new_state)`` with
    ``state = cat(c, h)`` along channels   (mcnet.py:259-294)."""

    def __init__(self, feature_size, num_features, forget_bias=1, activation=F.tanh, bias=True):
        super(ConvLstmCell, self).__init__()
        self.feature_size = feature_size
        self.num_features = num_features
        self.forget_bias = forget_bias
        self.activation = activation
        self.conv = nn.Conv2d(num_features * 2, num_features * 4, feature_size, padding=(feature_size - 1) // 2,
                              bias=bias)

    def gates(self, conv_output, state):
        """mcnet.py:287,290-293 as one kernel: returns new_state = cat(c', h')."""
        if not _is_tanh(self.activation):
            raise NotImplementedError("the fused gate kernel implements activation=tanh (the only one the "
                                      "reference's model registry uses, create_model.py:27-36)")
        return ops.ConvLstmGatesFunction.apply(conv_output.contiguous(), state.contiguous(), self.forget_bias)

    def forward(self, input, state):
        h = state[:, self.num_features:]
        x = torch.cat((input, h), dim=1)
        if x.is_cuda and self.conv.bias is not None:
            # bias-free cuDNN convolution + one vectorised in-place bias pass (layers.FusedSequential explains why)
            conv_output = F.conv2d(x, self.conv.weight, None, self.conv.stride, self.conv.padding)
            conv_output = ops.BiasActFunction.apply(conv_output.contiguous(), self.conv.bias, "none", 0.0)
        else:
            conv_output = self.conv(x)
        new_state = self.gates(conv_output, state)
        return new_state[:, self.num_features:], new_state


class MCNet(nn.Module):
    """MC-Net (Villegas et al.): ``forward(K, T, diff_in, xt) -> (pred, dyn, cont, res)`` lists of length T
    (mcnet.py:350-453)."""

    def __init__(self, gf_dim, c_dim, feature_size, forget_bias=1, activation=F.tanh, bias=True):
        super(MCNet, self).__init__()
        self.c_dim = c_dim
        self.gf_dim = gf_dim
        self.motion_enc = MotionEnc(gf_dim)
        self.conv_lstm_cell = ConvLstmCell(feature_size, 4 * gf_dim, forget_bias=forget_bias, activation=activation,
                                           bias=bias)
        self.content_enc = ContentEnc(c_dim, gf_dim)
        self.comb_layers = CombLayers(gf_dim)
        self.residual3 = Residual(gf_dim * 8, gf_dim * 4)
        self.residual2 = Residual(gf_dim * 4, gf_dim * 2)
        self.residual1 = Residual(gf_dim * 2, gf_dim * 1)
        self.dec_cnn = DecCnn(c_dim, gf_dim)

    batch_history = True  # see forward(); the CPU port of the reference switches it off

    def get_initial_conv_lstm_state(self, batch_size, image_size):
        ref = next(self.parameters())
        return torch.zeros(batch_size, 8 * self.gf_dim, image_size[0] // 8, image_size[1] // 8,
                           device=ref.device, dtype=ref.dtype)

    def _gray01(self, frame):
        x = inverse_transform(frame)
        return bgr2gray(x) if self.c_dim == 3 else x

    def forward(self, K, T, diff_in, xt):
        diffs = [d.squeeze(1) for d in torch.chunk(diff_in, diff_in.shape[1], dim=1)]
        image_size = xt.shape[2:4]
        state = self.get_initial_conv_lstm_state(xt.shape[0], image_size)

        # motion history (mcnet.py:405-409 encodes the K-1 known difference frames one by one; the encoder has no
        # state, so all but the last go through it as ONE batch -- the last stays separate because its skip
        # activations res_m are the ones the decoder uses -- and only the ConvLSTM walks them in order)
        if self.batch_history and K - 1 >= 3:
            N = xt.shape[0]
            enc_all, _ = self.motion_enc(torch.cat(diffs[:K - 2], 0))
            for enc_h in enc_all.view(K - 2, N, *enc_all.shape[1:]).unbind(0):
                h_dyn, state = self.conv_lstm_cell(enc_h, state)
            enc_h, res_m = self.motion_enc(diffs[K - 2])
            h_dyn, state = self.conv_lstm_cell(enc_h, state)
        else:
            for t in range(K - 1):
                enc_h, res_m = self.motion_enc(diffs[t])
                h_dyn, state = self.conv_lstm_cell(enc_h, state)

        pred, dyn, cont, res = [], [], [], []
        for t in range(T):
            if t > 0:
                enc_h, res_m = self.motion_enc(diffs[-1])
                h_dyn, state = self.conv_lstm_cell(enc_h, state)
            h_cont, res_c = self.content_enc(xt)
            h_tpl = self.comb_layers(h_dyn, h_cont)
            dyn.append(h_dyn)
            cont.append(h_cont)
            res_1 = self.residual1(res_m[0], res_c[0])
            res_2 = self.residual2(res_m[1], res_c[1])
            res_3 = self.residual3(res_m[2], res_c[2])
            res.append([res_1, res_2, res_3])
            x_hat = self.dec_cnn(h_tpl, res_1, res_2, res_3)
            # next motion input: difference of gray frames in [0, 1]   (mcnet.py:439-447), one kernel on the GPU
            if x_hat.is_cuda and self.c_dim in (1, 3):
                diffs.append(ops.GrayDiffPairFunction.apply(x_hat.contiguous(), xt.contiguous()))
            else:
                diffs.append(self._gray01(x_hat) - self._gray01(xt))
            xt = x_hat
            pred.append(x_hat.view(-1, self.c_dim, image_size[0], image_size[1]))
        return pred, dyn, cont, res


class MCNetFillInModel(nn.Module):
    """Forward-only baseline: ``forward(T, preceding, following) -> {'pred'}``   (mcnet.py:301-347)."""

    def __init__(self, gf_dim, c_dim, feature_size, forget_bias=1, activation=F.tanh, bias=True):
        super(MCNetFillInModel, self).__init__()
        self.c_dim = c_dim
        self.conv_lstm_state_size = 8 * gf_dim
        self.generator = MCNet(gf_dim, c_dim, feature_size, forget_bias=forget_bias, activation=activation, bias=bias)

    def forward(self, T, preceding_frames, following_frames):
        K = preceding_frames.size(1)
        xt = preceding_frames[:, -1]
        diff_in = gray_difference_frames(preceding_frames)
        forward_pred, _, _, _ = self.generator(K, T, diff_in, xt)
        return {'pred': torch.stack(forward_pred, dim=1)}


def gray_difference_frames(frames, reverse=False):
    """[B,K,C,H,W] in [-1,1] -> gray frames in [0,1] -> K-1 temporal differences   (tai.py:67-68); `reverse`: of the
    time-reversed clip (tai.py:71-74).  One kernel on the GPU (bit-identical to the elementwise chain below)."""
    if frames.is_cuda and frames.size(2) in (1, 3) and frames.size(1) > 1 and not (torch.is_grad_enabled() and frames.requires_grad):
        return ops.gray_difference_frames(frames.contiguous(), reverse)
    if reverse:
        frames = torch.flip(frames, dims=[1])
    x = inverse_transform(frames)
    gray = bgr2gray_batched(x) if frames.size(2) > 1 else x
    return gray[:, 1:] - gray[:, :-1]
