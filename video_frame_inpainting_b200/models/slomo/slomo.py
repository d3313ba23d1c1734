"""Super SloMo baseline with the warp / blend stages on sm_100a kernels.

Host-side mirror of the reference's ``src/models/slomo/slomo.py`` (same classes, attribute names and
return tuple).  The two U-Nets stay on cuDNN; the per-time-step glue changes:

* ``FlowWarper.forward`` (slomo.py:265-286) built a NumPy meshgrid on the host, copied it to the GPU on
  every call and ran ~8 elementwise kernels before ``F.grid_sample``; here it is one gather kernel
  (``flow_warp_forward_b200``) with the torch-0.3.1 sampling convention baked in (bilinear, zero padding,
  ``ix = ((g+1)/2)*(W-1)`` applied to ``g = 2*((x+u)/W - 0.5)`` -- so a zero flow samples ``x*(W-1)/W``,
  not ``x``, exactly like the reference);
* the T middle frames do not depend on each other (the U-Nets carry no state between them), so with
  ``batch_time`` (default) the per-t loop of slomo.py:307-340 becomes ONE pass over T*B samples: one kernel writes
  the flows of every t, both warps and the refinement network's input ``cat(I0, g0, F_t0, F_t1, g1, I1)`` in place
  (``slomo_interp_input_*``), the refinement U-Net runs once over the T*B batch, and one kernel does
  refine-clamp + two warps + visibility blend for every (t, b) straight into ``pred`` (``slomo_refine_blend_batched_*``).
  Both kernels have gather-only adjoints, so the training step takes the same route;
* the per-t formulation is kept (``batch_time = False``, or frames that require gradients): without autograd the
  flow combination + first two warps (slomo.py:312-316) and the refine-clamp + two warps + visibility blend
  (slomo.py:320-328) are one kernel each, with autograd the composed route through ``FlowWarper``.

Quirk kept on purpose: new frames are PREPENDED (slomo.py:332-340), so ``pred[:, 0]`` is the LAST middle
frame.
"""
from __future__ import division

import torch
import torch.nn as nn

from ... import ops
from ..layers import BilinearUp2, FusedSequential, MaxPool2


def _up2():
    return BilinearUp2()  # torch-0.3.1 bilinear mapping (align corners), one kernel of this library


def _stage(cin, cmid, cout, k, alpha, pool):
    layers = [MaxPool2()] if pool else []
    layers += [nn.Conv2d(cin, cmid, k, padding=k // 2), nn.LeakyReLU(alpha),
               nn.Conv2d(cmid, cout, k, padding=k // 2), nn.LeakyReLU(alpha)]
    return FusedSequential(*layers)


class Encoder(nn.Module):
    """Six-stage U-Net encoder; returns (bottleneck, [enc1..enc5])   (slomo.py:28-98)."""

    def __init__(self, gf_dim, input_dim, alpha=0.1):
        super(Encoder, self).__init__()
        g = gf_dim
        self.enc1 = _stage(input_dim, g, g, 7, alpha, pool=False)
        self.enc2 = _stage(g, g * 2, g * 2, 5, alpha, pool=True)
        self.enc3 = _stage(g * 2, g * 4, g * 4, 3, alpha, pool=True)
        self.enc4 = _stage(g * 4, g * 8, g * 8, 3, alpha, pool=True)
        self.enc5 = _stage(g * 8, g * 16, g * 16, 3, alpha, pool=True)
        self.enc6 = _stage(g * 16, g * 16, g * 16, 3, alpha, pool=True)

    def forward(self, input_imgs):
        feats = []
        x = input_imgs
        for stage in (self.enc1, self.enc2, self.enc3, self.enc4, self.enc5):
            x = stage(x)
            feats.append(x)
        return self.enc6(x), feats


class _Decoder(nn.Module):
    """Five (upsample, cat skip, 2 x conv) stages and a 1x1 output conv   (slomo.py:101-178, 181-262)."""

    def __init__(self, gf_dim, out_dim, alpha=0.1):
        super(_Decoder, self).__init__()
        g = gf_dim
        widths = [(g * 32, g * 16, g * 8), (g * 16, g * 8, g * 4), (g * 8, g * 4, g * 2), (g * 4, g * 2, g),
                  (g * 2, g, g)]
        for i, (cin, cmid, cout) in enumerate(widths, start=1):
            setattr(self, "upsample%d" % i, _up2())
            setattr(self, "dec%d" % i, _stage(cin, cmid, cout, 3, alpha, pool=False))
        self.output = nn.Conv2d(g, out_dim, 1)

    def decode(self, encoded_input, res_in):
        x = encoded_input
        for i in range(1, 6):
            x = getattr(self, "upsample%d" % i)(x)
            x = getattr(self, "dec%d" % i)(torch.cat((x, res_in[-i]), 1))
        return self.output(x)


class ComputeDecoder(_Decoder):
    """Flow computation head: tanh on all 4 channels (F_0_1, F_1_0)."""

    def __init__(self, gf_dim, out_dim, alpha=0.1):
        super(ComputeDecoder, self).__init__(gf_dim, out_dim, alpha)
        self.tanh = nn.Tanh()

    def forward(self, encoded_input, res_in):
        return self.tanh(self.decode(encoded_input, res_in))


class RefineDecoder(_Decoder):
    """Refinement head: (tanh dF_t0, tanh dF_t1, sigmoid V_t0)   (slomo.py:255-262)."""

    def __init__(self, gf_dim, out_dim, alpha=0.1):
        super(RefineDecoder, self).__init__(gf_dim, out_dim, alpha)
        self.sigmoid = nn.Sigmoid()
        self.tanh = nn.Tanh()

    def forward(self, encoded_input, res_in):
        delta_F_t_0, delta_F_t_1, V_t_0 = torch.split(self.decode(encoded_input, res_in), 2, dim=1)
        return self.tanh(delta_F_t_0), self.tanh(delta_F_t_1), self.sigmoid(V_t_0)


class FlowWarper(nn.Module):
    """``forward(img[B,C,H,W], uv[B,2,H,W])`` -> backward-warped image   (slomo.py:265-286)."""

    def forward(self, img, uv):
        return ops.FlowWarpFunction.apply(img.contiguous(), uv.contiguous())


class SloMo(nn.Module):
    """``forward(T, I0, I1)`` -> (predictions[B,T,C,H,W], F_0_1, F_1_0, F_t_0_collector, F_t_1_collector)
    (slomo.py:289-342)."""

    def __init__(self, gf_dim, c_input_dim):
        super(SloMo, self).__init__()
        self.c_input_dim = c_input_dim
        self.compute_enc = Encoder(gf_dim, 2 * c_input_dim)
        self.compute_dec = ComputeDecoder(gf_dim, 4)
        self.flow_warper = FlowWarper()
        self.refine_enc = Encoder(gf_dim, 4 * c_input_dim + 4)
        self.refine_dec = RefineDecoder(gf_dim, 5)
        self.batch_time = True   # all T middle frames as one batch (same arithmetic per sample)

    def intermediate_flows_and_warps(self, I0, I1, F_0_1, F_1_0, t, differentiable):
        """slomo.py:312-316 -> (F_t_0, F_t_1, g_I0_F_t_0, g_I1_F_t_1); one kernel when no gradient is needed."""
        if not differentiable:
            return ops.slomo_flow_combine_warp(I0, I1, F_0_1, F_1_0, t)
        F_t_0 = -(1 - t) * t * F_0_1 + t ** 2 * F_1_0
        F_t_1 = (1 - t) * (1 - t) * F_0_1 - t * (1 - t) * F_1_0
        return F_t_0, F_t_1, self.flow_warper(I0, F_t_0), self.flow_warper(I1, F_t_1)

    def refine_and_blend(self, I0, I1, F_t_0, F_t_1, delta_F_t_0, delta_F_t_1, V_t_0, t, differentiable):
        """slomo.py:320-328 -> interpolated frame; one kernel when no gradient is needed."""
        if not differentiable:
            return ops.slomo_refine_blend(I0, I1, F_t_0, F_t_1, delta_F_t_0.contiguous(), delta_F_t_1.contiguous(),
                                          V_t_0.contiguous(), t)
        F_t_0_refine = torch.clamp(delta_F_t_0 + F_t_0, min=-1, max=1)
        F_t_1_refine = torch.clamp(delta_F_t_1 + F_t_1, min=-1, max=1)
        V_t_1 = 1 - V_t_0
        g0 = self.flow_warper(I0, F_t_0_refine)
        g1 = self.flow_warper(I1, F_t_1_refine)
        normalization = (1 - t) * V_t_0 + t * V_t_1
        return ((1 - t) * V_t_0 * g0 + t * V_t_1 * g1) / normalization

    def forward(self, T, I0, I1):
        I0, I1 = I0.contiguous(), I1.contiguous()
        img = torch.cat((I0, I1), 1)
        flows = self.compute_dec(*self.compute_enc(img))
        F_0_1 = flows[:, :2].contiguous()
        F_1_0 = flows[:, 2:].contiguous()
        differentiable = torch.is_grad_enabled() and any(x.requires_grad for x in (F_0_1, I0, I1))
        frames_need_grad = torch.is_grad_enabled() and (I0.requires_grad or I1.requires_grad)
        if self.batch_time and I0.is_cuda and not frames_need_grad and T <= 16:
            # one pass over the T*B samples (sample n = t*B + b); collectors and pred come out in the reference's
            # reversed time order (slomo.py:332-340)
            interp_input, F_t_0_collector, F_t_1_collector = ops.SlomoInterpInputFunction.apply(I0, I1, F_0_1, F_1_0, T)
            # the warped frames g(I0, F_t_0), g(I1, F_t_1) of every t are channels of this tensor: the training
            # environment's warping loss reads them here instead of warping again (environments.py:584-586)
            self.last_interp_input = interp_input if self.training else None
            delta_F_t_0, delta_F_t_1, V_t_0 = self.refine_dec(*self.refine_enc(interp_input))
            pred = ops.SlomoRefineBlendFunction.apply(I0, I1, F_t_0_collector, F_t_1_collector,
                                                      delta_F_t_0.contiguous(), delta_F_t_1.contiguous(),
                                                      V_t_0.contiguous(), T)
            return pred, F_0_1, F_1_0, F_t_0_collector, F_t_1_collector
        self.last_interp_input = None
        preds, ft0s, ft1s = [], [], []
        for t_ in range(T):
            t = (t_ + 1) / (T + 1)
            F_t_0, F_t_1, g_I0_F_t_0, g_I1_F_t_1 = self.intermediate_flows_and_warps(I0, I1, F_0_1, F_1_0, t,
                                                                                    differentiable)
            interp_input = torch.cat((I0, g_I0_F_t_0, F_t_0, F_t_1, g_I1_F_t_1, I1), 1)
            delta_F_t_0, delta_F_t_1, V_t_0 = self.refine_dec(*self.refine_enc(interp_input))
            interp_image = self.refine_and_blend(I0, I1, F_t_0, F_t_1, delta_F_t_0, delta_F_t_1, V_t_0, t,
                                                 differentiable)
            # prepend: the reference's collectors end up in reverse time order (slomo.py:332-340)
            preds.insert(0, interp_image)
            ft0s.insert(0, F_t_0)
            ft1s.insert(0, F_t_1)
        return torch.stack(preds, 1), F_0_1, F_1_0, torch.stack(ft0s, 1), torch.stack(ft1s, 1)


class SloMoFillInModel(nn.Module):
    """Uses only the last preceding and the first following frame   (slomo.py:345-371)."""

    def __init__(self, gf_dim=32, c_input_dim=3):
        super(SloMoFillInModel, self).__init__()
        self.generator = SloMo(gf_dim, c_input_dim)

    def forward(self, T, preceding_frames, following_frames):
        pred, F_0_1, F_1_0, F_t_0_collector, F_t_1_collector = self.generator(
            T, preceding_frames[:, -1].contiguous(), following_frames[:, 0].contiguous())
        return {'pred': pred, 'F_0_1': F_0_1, 'F_1_0': F_1_0,
                'F_t_0_collector': F_t_0_collector, 'F_t_1_collector': F_t_1_collector}
