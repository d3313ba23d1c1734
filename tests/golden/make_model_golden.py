"""Generates tests/golden/tai_model_ref.npz by IMPORTING the reference's own model classes
(/root/reference/src/models/{tai,mcnet}/*.py, unmodified) and running them on CPU tensors.

Run in the build container only (the reference is not present on the GPU box):

    python tests/golden/make_model_golden.py

The reference is Python-2 / torch-0.3.1 code with no CPU implementation of its operator, so the import needs
shims.  Each one is listed here because it bounds what the fixture pins:

  * ``xrange`` -> ``range``; a ``Queue`` module (util.py:5); stub modules for matplotlib / tensorboardX / imageio /
    skimage (plotting and logging, never executed here); a stub ``_ext.cunnex`` (the cffi extension).
  * ``src.separable_convolution.SeparableConvolution.SeparableConvolution`` is replaced by the C port of the
    reference kernels (oracle/reference_model.py: CpuSeparableConvolution, pinned separately against the reference's
    CUDA kernels) -- the reference operator raises NotImplementedError on CPU tensors (SeparableConvolution.py:48-49).
  * ``Tensor.cuda`` is the identity (tai.py:72,216 and mcnet.py:386 move helper tensors to the GPU unconditionally).
  * Python-2 integer division: ``nn.Conv2d`` paddings and ``torch.zeros`` sizes are coerced to int
    (mcnet.py:278 ``(feature_size - 1) / 2``, mcnet.py:384 ``image_size[0]/8``).
  * ``nn.Upsample(mode='bilinear')`` and ``F.grid_sample`` are evaluated with ``align_corners=True``: torch 0.3.1's
    only mapping.  (This one is an assumption about the un-vendored library, see oracle/oracle.py's header: the
    fixture pins the models' composition, not that mapping.)

What the fixture therefore pins, through the reference's own code: the module tree and every state_dict key and
shape (checkpoint compatibility), the order of the forward / backward streams and the time reversal, the residual
merging, the time-ratio injection, the ConvLSTM gate chain, the zero-insertion unpooling chain, the gray /
difference-frame prologue and the 0.5 / 0.5 blend.
"""
import builtins
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def install_shims():
    import queue
    builtins.xrange = range
    q = types.ModuleType('Queue')
    q.Queue = queue.Queue
    sys.modules['Queue'] = q
    mpl = types.ModuleType('matplotlib')
    mpl.use = lambda *a, **k: None
    sys.modules['matplotlib'] = mpl
    sys.modules['matplotlib.pyplot'] = types.ModuleType('matplotlib.pyplot')
    for name in ('tensorboardX', 'imageio', 'skimage', 'skimage.measure'):
        sys.modules.setdefault(name, types.ModuleType(name))
    ext = types.ModuleType('_ext')
    ext.cunnex = types.ModuleType('_ext.cunnex')
    sys.modules['_ext'] = ext
    sys.modules['_ext.cunnex'] = ext.cunnex
    torch.Tensor.cuda = lambda self, *a, **k: self

    conv_init = nn.Conv2d.__init__

    def conv_init_int(self, *args, **kwargs):
        if 'padding' in kwargs and isinstance(kwargs['padding'], float):
            kwargs['padding'] = int(kwargs['padding'])
        conv_init(self, *args, **kwargs)
    nn.Conv2d.__init__ = conv_init_int

    zeros = torch.zeros

    def zeros_int(*size, **kwargs):
        return zeros(*[int(s) if isinstance(s, float) else s for s in size], **kwargs)
    torch.zeros = zeros_int

    def upsample_forward(self, x):
        assert self.mode == 'bilinear'
        return F.interpolate(x, scale_factor=self.scale_factor, mode='bilinear', align_corners=True)
    nn.Upsample.forward = upsample_forward

    grid_sample = F.grid_sample

    def grid_sample_031(img, grid, *a, **k):   # torch 0.3.1: bilinear, zero padding, ix = ((g + 1) / 2) * (W - 1)
        return grid_sample(img, grid, mode='bilinear', padding_mode='zeros', align_corners=True)
    F.grid_sample = grid_sample_031


def main():
    install_shims()
    sys.path.insert(0, '/root/reference')
    from oracle.reference_model import CpuSeparableConvolution
    import src.separable_convolution.SeparableConvolution as ref_op
    ref_op.SeparableConvolution = CpuSeparableConvolution
    import src.models.tai.tai as ref_tai
    import src.util.util as ref_util

    out = {}
    cases = [dict(gf_dim=4, c_dim=1, feature_size=3, ks=13, num_block=5, kf_dim=4, K=3, T=2, F_=3, H=32, W=32, B=2),
             dict(gf_dim=4, c_dim=3, feature_size=3, ks=5, num_block=4, kf_dim=2, K=2, T=3, F_=3, H=32, W=48, B=1)]
    for ci, c in enumerate(cases):
        torch.manual_seed(100 + ci)
        model = ref_tai.TAIFillInModel(c['gf_dim'], c['c_dim'], c['feature_size'], c['ks'], num_block=c['num_block'],
                                       kf_dim=c['kf_dim'])
        model.apply(ref_util.weights_init)
        # xavier weights make this small network's outputs vanish; scale the biases up so that every branch carries signal
        g = torch.Generator().manual_seed(200 + ci)
        for name, p in model.named_parameters():
            if name.endswith('bias'):
                p.data.uniform_(-0.1, 0.1, generator=g)
        pre = torch.rand(c['B'], c['K'], c['c_dim'], c['H'], c['W'], generator=g) * 2 - 1
        fol = torch.rand(c['B'], c['F_'], c['c_dim'], c['H'], c['W'], generator=g) * 2 - 1
        res = model(c['T'], pre, fol)
        loss = res['pred'].pow(2).mean() + res['pred_forward'].mean() + res['pred_backward'].pow(2).mean()
        loss.backward()
        tag = 'c%d_' % ci
        for k, v in c.items():
            out[tag + 'cfg_' + k] = np.int64(v)
        out[tag + 'pre'] = pre.numpy()
        out[tag + 'fol'] = fol.numpy()
        for k, v in res.items():
            out[tag + 'out_' + k] = v.detach().numpy()
        names = []
        for name, v in model.state_dict().items():
            names.append(name)
            out[tag + 'sd_' + name] = v.numpy()
        out[tag + 'sd_names'] = np.array(names)
        # gradients of a few parameters spread over the network (first / last layers of each sub-network)
        picks = [n for n, _ in model.named_parameters()]
        picks = [picks[0], picks[len(picks) // 3], picks[2 * len(picks) // 3], picks[-1]]
        params = dict(model.named_parameters())
        for n in picks:
            out[tag + 'grad_' + n] = params[n].grad.numpy()
        out[tag + 'grad_names'] = np.array(picks)
        print('case', ci, 'params', sum(p.numel() for p in model.parameters()), 'pred rms', float(res['pred'].pow(2).mean().sqrt()))
    # bi-TWI ablation (src/models/twi/twi.py: time-weighted blend, no time-ratio plane) and the forward-only MC-Net
    # baseline (src/models/mcnet/mcnet.py: MCNetFillInModel)
    import src.models.twi.twi as ref_twi
    import src.models.mcnet.mcnet as ref_mcnet
    torch.manual_seed(600)
    extra = {'twi': ref_twi.TimeWeightedInterpolationFillInModel(4, 1, 3, 5, num_block=5, kf_dim=2),
             'mcnet': ref_mcnet.MCNetFillInModel(4, 3, 3)}
    g = torch.Generator().manual_seed(601)
    for tag, em in extra.items():
        em.apply(ref_util.weights_init)
        for name, p in em.named_parameters():
            if name.endswith('bias'):
                p.data.uniform_(-0.1, 0.1, generator=g)
        c = 1 if tag == 'twi' else 3
        pre = torch.rand(2, 3, c, 32, 32, generator=g) * 2 - 1
        fol = torch.rand(2, 3, c, 32, 32, generator=g) * 2 - 1
        res = em(3, pre, fol)
        res['pred'].pow(2).mean().backward()
        out[tag + '_pre'], out[tag + '_fol'] = pre.numpy(), fol.numpy()
        for k, v in res.items():
            out[tag + '_out_' + k] = v.detach().numpy()
        names = []
        for name, v in em.state_dict().items():
            names.append(name)
            out[tag + '_sd_' + name] = v.numpy()
        out[tag + '_sd_names'] = np.array(names)
        first = [n for n, _ in em.named_parameters()][0]
        out[tag + '_grad_name'] = np.array([first])
        out[tag + '_grad'] = dict(em.named_parameters())[first].grad.numpy()
        print(tag, 'params', sum(p.numel() for p in em.parameters()), 'outputs', sorted(res))

    # Super SloMo baseline (src/models/slomo/slomo.py): flow combination, warps, refinement, visibility blend, and
    # the reversed time order in which the reference concatenates its predictions (slomo.py:331-340)
    import src.models.slomo.slomo as ref_slomo
    torch.manual_seed(300)
    sm = ref_slomo.SloMoFillInModel(gf_dim=2, c_input_dim=3)
    sm.apply(ref_util.weights_init)
    g = torch.Generator().manual_seed(301)
    for name, p in sm.named_parameters():
        if name.endswith('bias'):
            p.data.uniform_(-0.1, 0.1, generator=g)
        else:
            p.data.mul_(1.75)             # flows of a fraction of a pixel instead of ~0 (x3 saturates the flow heads: chaotic)
    # smooth frames (an upsampled 8 x 16 noise grid): a warp of white noise would turn 1e-4 of flow difference between
    # two convolution implementations into 1e-2 of image difference
    def smooth():
        return F.interpolate(torch.rand(2, 3, 8, 16, generator=g) * 2 - 1, size=(32, 64), mode='bilinear',
                             align_corners=True).unsqueeze(0)
    pre, fol = smooth(), smooth()
    res = sm(3, pre, fol)
    res['pred'].pow(2).mean().backward()
    out['s_pre'], out['s_fol'] = pre.numpy(), fol.numpy()
    for k, v in res.items():
        out['s_out_' + k] = v.detach().numpy()
    names = []
    for name, v in sm.state_dict().items():
        names.append(name)
        out['s_sd_' + name] = v.numpy()
    out['s_sd_names'] = np.array(names)
    first = [n for n, _ in sm.named_parameters()][0]
    out['s_grad_name'] = np.array([first])
    out['s_grad'] = dict(sm.named_parameters())[first].grad.numpy()
    print('slomo params', sum(p.numel() for p in sm.parameters()), 'pred rms', float(res['pred'].detach().pow(2).mean().sqrt()),
          'flow rms', float(res['F_0_1'].detach().pow(2).mean().sqrt()))
    # Spectral-norm discriminator (src/discriminators/SNDiscriminator.py): window slicing, one power-iteration update
    # per SN layer and window with the in-place division of weight.data, u carried between calls.  The reference
    # draws the initial u from the global RNG on first use; the fixture presets it on both sides (both
    # implementations take a given u as it is, SNDiscriminator.py:16-20).  Shim: torch >= 1.x's _ConvNd.__init__
    # takes a padding_mode argument that SNConv2d (SNDiscriminator.py:60-61) does not pass.
    from torch.nn.modules import conv as conv_mod
    convnd_init = conv_mod._ConvNd.__init__

    def convnd_init_compat(self, *args, **kwargs):
        if len(args) == 10 and 'padding_mode' not in kwargs:
            args = args + ('zeros',)
        convnd_init(self, *args, **kwargs)
    conv_mod._ConvNd.__init__ = convnd_init_compat
    import src.discriminators.SNDiscriminator as ref_disc
    torch.manual_seed(400)
    disc = ref_disc.SNDiscriminator((32, 32), 1, 3, 4, 3)
    disc.apply(ref_util.weights_init)
    g = torch.Generator().manual_seed(401)
    sn_layers = [m for m in disc.modules() if hasattr(m, 'Ip')]
    for i, m in enumerate(sn_layers):
        m.u = torch.randn(1, m.weight.size(0), generator=g)
        out['d_u%d' % i] = m.u.numpy().copy()
    names = []
    for name, v in disc.state_dict().items():
        names.append(name)
        out['d_sd_' + name] = v.numpy().copy()          # weights BEFORE the first call (forward normalises in place)
    out['d_sd_names'] = np.array(names)
    video = torch.rand(2, 6, 1, 32, 32, generator=g) * 2 - 1
    out['d_video'] = video.numpy()
    out['d_logits_call1'] = disc(video).detach().numpy()
    logits2 = disc(video)
    out['d_logits_call2'] = logits2.detach().numpy()
    logits2.sum().backward()
    out['d_weight0_after'] = disc.conv_layers[0].weight.detach().numpy().copy()
    out['d_grad_weight0'] = disc.conv_layers[0].weight.grad.numpy().copy()
    out['d_grad_linear'] = disc.linear_layer.weight.grad.numpy().copy()
    print('discriminator params', sum(p.numel() for p in disc.parameters()), 'logits', out['d_logits_call2'].shape)
    out['n'] = np.int64(len(cases))
    path = os.path.join(HERE, 'tai_model_ref.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path) // 1024, 'KB')


if __name__ == '__main__':
    main()
