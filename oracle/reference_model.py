"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference MODEL on the TAI path.

The reference has no CPU implementation (its operator raises NotImplementedError on CPU tensors,
SeparableConvolution.py:48-49; options.py:61 asserts a GPU) and cannot be imported under Python 3 /
torch 2.  The CPU baseline that stands in for it is
therefore a port: the same module tree as the product (so the convolution stack is identical and runs
on torch's CPU kernels), with every hot-path operator replaced by the reference's own formulation:

* separable convolution  -> the literal FP32 C port of kernel.cu:19-162 (oracle/sepconv_oracle.c),
  called once per stream after an explicit ReplicationPad2d, followed by the blend as three
  elementwise ops (tai.py:229-236, 105);
* ConvLSTM gates          -> the chunk / sigmoid / tanh / cat chain of mcnet.py:287-293;
* FlowWarper               -> host meshgrid + F.grid_sample with the torch-0.3.1 mapping (slomo.py:265-286);
* DecCnn unpool + add      -> the permute / cat / clone().zero_() chain of mcnet.py:240-256 and the add of 234-236;
* nn.Upsample (bilinear)   -> F.interpolate(align_corners=True), the torch-0.3.1 mapping;
* nn.MaxPool2d(2)          -> F.max_pool2d (the library op the reference calls).

Parity status: PINNED for the TAI model composition -- tests/golden/tai_model_ref.npz holds a state_dict, inputs,
outputs and gradients produced by the reference's own TAIFillInModel / MCNet / TAI classes (imported unmodified by
tests/golden/make_model_golden.py; shims listed there) and tests/test_models_cpu.py checks this port against it
(1e-5 outputs, 1e-4 gradients, strict state_dict load); the same for SloMoFillInModel (slomo.py).  The training step
(environments.py is a SyntaxError under Python 3) is restated from the source only; the SN discriminator it uses
(the product's torch module) is pinned on the reference's own SNDiscriminator classes by the same fixture.

Used by bench.py (``cpu_baseline`` and ``--impl reference``) and by tests as an end-to-end checker of
the GPU model.  Never imported by the product package.
"""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import oracle as O
from video_frame_inpainting_b200.discriminators.SNDiscriminator import SNDiscriminator
from video_frame_inpainting_b200.losses.losses import GDL
from video_frame_inpainting_b200.models.layers import BilinearUp2, MaxPool2
from video_frame_inpainting_b200.models.mcnet.mcnet import ConvLstmCell, DecCnn
from video_frame_inpainting_b200.models.slomo.slomo import FlowWarper, SloMo
from video_frame_inpainting_b200.models.tai.tai import TAI
from video_frame_inpainting_b200.util.util import inverse_transform, weights_init


class CpuSeparableConvolution(torch.autograd.Function):
    """SeparableConvolution.py:6-92 with the kernels replaced by their C port (FP32, reference loop order)."""

    @staticmethod
    def forward(ctx, input, vertical, horizontal, ks=51):
        ctx.save_for_backward(input, vertical, horizontal)
        ctx.constant = ks
        out = O.sepconv_forward(input.numpy(), vertical.numpy(), horizontal.numpy(), ks, dtype=np.float32)
        return torch.from_numpy(out)

    @staticmethod
    def backward(ctx, grad_output):
        input, vertical, horizontal = ctx.saved_tensors
        ks = ctx.constant
        go = grad_output.contiguous().numpy()
        gi = O.sepconv_grad_input(go, vertical.numpy(), horizontal.numpy(), ks, dtype=np.float32)
        gv = O.sepconv_grad_vertical(go, input.numpy(), horizontal.numpy(), ks, dtype=np.float32)
        gh = O.sepconv_grad_horizontal(go, input.numpy(), vertical.numpy(), ks, dtype=np.float32)
        return torch.from_numpy(gi), torch.from_numpy(gv), torch.from_numpy(gh), None


class CpuConvLstmCell(ConvLstmCell):
    def gates(self, conv_output, state):
        c, _ = torch.chunk(state, 2, dim=1)
        i, j, f, o = torch.chunk(conv_output, 4, dim=1)
        new_c = c * torch.sigmoid(f + self.forget_bias) + torch.sigmoid(i) * torch.tanh(j)
        new_h = torch.tanh(new_c) * torch.sigmoid(o)
        return torch.cat((new_c, new_h), dim=1)


class CpuDecCnn(DecCnn):
    """fixed_unpooling as the reference spells it (mcnet.py:240-256) followed by the add (234-236)."""

    def unpool_add(self, x, res):
        x = x.permute(0, 2, 3, 1)
        out = torch.cat((x, x.clone().zero_()), dim=3)
        out = torch.cat((out, out.clone().zero_()), dim=2)
        return out.view(x.size(0), 2 * x.size(1), 2 * x.size(2), x.size(3)).permute(0, 3, 1, 2) + res


class CpuBilinearUp2(BilinearUp2):
    """nn.Upsample(scale_factor=2, mode='bilinear') with the torch-0.3.1 mapping (align_corners=True today)."""

    def forward(self, x):
        return F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=True)


class CpuTAIMixin(object):
    """Reference formulation of the filter-and-blend tail for TAI and its TWI subclass."""

    def apply_maps(self, variableInput1, variableInput2, v1, h1, v2, h2, a=0.5, b=0.5):
        apply = CpuSeparableConvolution.apply
        dot1 = apply(self.modulePad(variableInput1).contiguous(), v1.contiguous(), h1.contiguous(), self.ks)
        dot2 = apply(self.modulePad(variableInput2).contiguous(), v2.contiguous(), h2.contiguous(), self.ks)
        return a * dot1 + b * dot2, dot1, dot2


class CpuFlowWarper(FlowWarper):
    def forward(self, img, uv):
        H, W = int(img.shape[-2]), int(img.shape[-1])
        gx, gy = np.meshgrid(np.arange(0, W), np.arange(0, H))
        X = torch.Tensor(gx).unsqueeze(0) + uv[:, 0]
        Y = torch.Tensor(gy).unsqueeze(0) + uv[:, 1]
        grid = torch.stack((2 * (X / W - 0.5), 2 * (Y / H - 0.5)), dim=3)
        return F.grid_sample(img, grid, mode='bilinear', padding_mode='zeros', align_corners=True)


class CpuMaxPool2(MaxPool2):
    """nn.MaxPool2d(2) as the reference calls it (mcnet.py:28-45; slomo.py:47-85)."""

    def forward(self, x):
        return F.max_pool2d(x, 2)


class CpuSloMo(SloMo):
    """Always the composed formulation of slomo.py:311-340 (no fused kernels)."""

    def intermediate_flows_and_warps(self, I0, I1, F_0_1, F_1_0, t, differentiable):
        return SloMo.intermediate_flows_and_warps(self, I0, I1, F_0_1, F_1_0, t, True)

    def refine_and_blend(self, I0, I1, F_t_0, F_t_1, delta_F_t_0, delta_F_t_1, V_t_0, t, differentiable):
        return SloMo.refine_and_blend(self, I0, I1, F_t_0, F_t_1, delta_F_t_0, delta_F_t_1, V_t_0, t, True)


def to_cpu_reference(model):
    """Re-class the hot-path modules of a (CPU-resident) product model to their reference formulations."""
    for m in model.modules():
        if hasattr(m, 'batch_streams'):
            m.batch_streams = False      # the reference runs the two MC-Net streams back to back (tai.py:77-84)
            m.batch_time = False         # ... and the kernel network once per middle frame (tai.py:91-105)
        if hasattr(m, 'batch_history'):
            m.batch_history = False      # one motion-encoder call per known difference frame (mcnet.py:405-409)
        if type(m) is ConvLstmCell:
            m.__class__ = CpuConvLstmCell
        elif type(m) is FlowWarper:
            m.__class__ = CpuFlowWarper
        elif type(m) is SloMo:
            m.__class__ = CpuSloMo
            m.batch_time = False         # the reference's per-t loop (slomo.py:307-340)
        elif type(m) is DecCnn:
            m.__class__ = CpuDecCnn
        elif type(m) is BilinearUp2:
            m.__class__ = CpuBilinearUp2
        elif type(m) is MaxPool2:
            m.__class__ = CpuMaxPool2
        elif isinstance(m, TAI) and not isinstance(m, CpuTAIMixin):
            m.__class__ = type('Cpu' + type(m).__name__, (CpuTAIMixin, type(m)), {})
            m.separableConvolution = CpuSeparableConvolution.apply
    return model


class CpuTAITrainingStep(object):
    """The TAI training step (environments.py:222-228, 326-379, 429-453) on CPU tensors: same losses,
    optimisers and update order as video_frame_inpainting_b200.environments.TAITrainingEnvironment.
    Pinned: tests/golden/tai_step_ref.npz holds one step of the reference's own TAITrainingEnvironment
    (tests/golden/make_step_golden.py); tests/test_models_cpu.py checks every loss term and the updated
    parameters of this port against it."""

    def __init__(self, generator, image_size, c_dim, K, T, F_, alpha=1.0, beta=0.02, lr=1e-4, beta1=0.5, df_dim=64,
                 Ip=3, disc_t=3):
        self.generator = to_cpu_reference(generator)
        self.generator.apply(weights_init)
        self.discriminator = SNDiscriminator(image_size, c_dim, disc_t, df_dim, Ip)
        self.discriminator.apply(weights_init)
        self.K, self.T, self.F, self.disc_t = K, T, F_, disc_t
        self.alpha, self.beta = alpha, beta
        self.mse, self.gdl, self.bce = torch.nn.MSELoss(), GDL(), torch.nn.BCEWithLogitsLoss()
        self.optimizer_G = torch.optim.Adam(self.generator.parameters(), lr=lr, betas=(beta1, 0.999))
        self.optimizer_D = torch.optim.Adam(self.discriminator.parameters(), lr=lr, betas=(beta1, 0.999))

    @staticmethod
    def _tm01(frames):
        _, _, c, H, W = frames.shape
        return inverse_transform(frames.permute(1, 0, 2, 3, 4).contiguous().view(-1, c, H, W))

    def _fake_labels(self):
        n = self.K + self.T + self.F - self.disc_t + 1
        ones_P, ones_F = max(0, self.K - self.disc_t + 1), max(0, self.F - self.disc_t + 1)
        labels = torch.zeros(n)
        labels[:ones_P] = 1
        if ones_F > 0:
            labels[n - ones_F:] = 1
        return labels

    def step(self, preceding, following, gt_middle):
        out = self.generator(self.T, preceding, following)
        gt = self._tm01(gt_middle)
        self.optimizer_G.zero_grad()
        loss = 0
        self.terms = {}
        for key, tag in (('pred', ''), ('pred_forward', '_forward'), ('pred_backward', '_backward')):
            x = self._tm01(out[key])
            lp, gdl = self.mse(x, gt), self.gdl(x, gt)
            self.terms['Lp' + tag], self.terms['gdl' + tag] = float(lp.detach()), float(gdl.detach())
            loss = loss + self.alpha * (lp + gdl)
        video = torch.cat([preceding, out['pred'], following], dim=1)
        h = self.discriminator(video)
        l_gan = self.bce(h, torch.ones_like(h))
        loss = loss + self.beta * l_gan
        loss.backward()
        self.optimizer_G.step()
        self.optimizer_D.zero_grad()
        h = self.discriminator(video.detach())
        labels = self._fake_labels().view(1, -1).expand(h.size(0), -1)
        h_real = self.discriminator(torch.cat([preceding, gt_middle, following], dim=1))
        l_fake, l_real = self.bce(h, labels), self.bce(h_real, torch.ones_like(h_real))
        loss_d = l_fake + l_real
        loss_d.backward()
        self.optimizer_D.step()
        self.terms.update(L_GAN=float(l_gan.detach()), loss_G=float(loss.detach()), loss_d_fake=float(l_fake.detach()),
                          loss_d_real=float(l_real.detach()), loss_D=float(loss_d.detach()))
        return float(loss.detach()), float(loss_d.detach())
