"""Generates tests/golden/tai_step_ref.npz: ONE training step of the reference's own ``TAITrainingEnvironment``
(src/environments/environments.py) on CPU tensors -- generator + spectral-norm discriminator, the seven loss terms,
both Adam updates.

    python tests/golden/make_step_golden.py          (build container only)

environments.py is not importable under Python 3: ``.cuda(async=True)`` (lines 94-98, 169-171) is a SyntaxError.
The module source is therefore read from /root/reference, the token ``async=True`` replaced by
``non_blocking=True`` (the torch >= 0.4 spelling of the same argument) and executed as
``src.environments.environments``; nothing else in it is changed.  Its imports of model families outside the TAI
path (self-attention, optical-flow, bi-SA, bi-TWA, TW_P_F) are satisfied by empty stand-in classes when the real
modules do not import under Python 3.  All other shims are those of make_model_golden.py (py2 names, Tensor.cuda
as identity, the C port for the CUDA-only operator, align_corners=True, _ConvNd signature).
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


def load_reference_environments():
    from make_model_golden import install_shims
    install_shims()
    from torch.nn.modules import conv as conv_mod
    convnd_init = conv_mod._ConvNd.__init__

    def convnd_init_compat(self, *args, **kwargs):
        if len(args) == 10 and 'padding_mode' not in kwargs:
            args = args + ('zeros',)
        convnd_init(self, *args, **kwargs)
    conv_mod._ConvNd.__init__ = convnd_init_compat
    sys.path.insert(0, '/root/reference')
    from oracle.reference_model import CpuSeparableConvolution
    import src.separable_convolution.SeparableConvolution as ref_op
    ref_op.SeparableConvolution = CpuSeparableConvolution
    wanted = {'src.models.self_attention.self_attention': ['BaseSCTSkipConFillInModel'],
              'src.models.optical_flow_fill_in.OFFillInModel': ['OFFillInModel'],
              'src.models.bi_sa.bi_sa': ['BidirectionalSimpleAverageFillInModel'],
              'src.models.bi_twa.bi_twa': ['BidirectionalTimeWeightedAverageFillInModel'],
              'src.models.tw_p_f.tw_p_f': ['TimeWeightedPFFillInModel']}
    stubbed = []
    for name, classes in wanted.items():
        try:
            importlib.import_module(name)
        except Exception:
            m = types.ModuleType(name)
            for c in classes:
                setattr(m, c, type(c, (), {}))
            sys.modules[name] = m
            stubbed.append(name)
    import src.environments  # noqa: F401  (the package; its __init__ is empty)
    path = '/root/reference/src/environments/environments.py'
    text = open(path).read()
    assert text.count('async=True') == 6
    mod = types.ModuleType('src.environments.environments')
    mod.__package__ = 'src.environments'
    mod.__file__ = path
    sys.modules['src.environments.environments'] = mod
    exec(compile(text.replace('async=True', 'non_blocking=True'), path, 'exec'), mod.__dict__)
    return mod, stubbed


def main():
    env_mod, stubbed = load_reference_environments()
    print('stand-ins for', stubbed)
    import src.models.tai.tai as ref_tai
    cfg = dict(gf_dim=4, c_dim=1, feature_size=3, ks=5, num_block=5, kf_dim=2, K=3, T=2, F_=3, H=32, W=32, B=2,
               df_dim=4, Ip=3, disc_t=3)
    hp = dict(alpha=1.0, beta=0.02, lr=1e-4, beta1=0.5)
    torch.manual_seed(500)
    model = ref_tai.TAIFillInModel(cfg['gf_dim'], cfg['c_dim'], cfg['feature_size'], cfg['ks'], num_block=cfg['num_block'],
                                   kf_dim=cfg['kf_dim'])
    env = env_mod.TAITrainingEnvironment(model, '/tmp/tai_ref_env', 'golden', (cfg['H'], cfg['W']), cfg['c_dim'],
                                         hp['alpha'], hp['beta'], hp['lr'], hp['beta1'], cfg['df_dim'], cfg['Ip'],
                                         cfg['disc_t'], cfg['K'], cfg['T'], cfg['F_'], (0, 0))
    g = torch.Generator().manual_seed(501)
    for name, p in env.generator.named_parameters():          # biases away from zero: every branch carries signal
        if name.endswith('bias'):
            p.data.uniform_(-0.1, 0.1, generator=g)
    out = {}
    for k, v in {**cfg, **hp}.items():
        out['cfg_' + k] = np.float64(v)
    sn_layers = [m for m in env.discriminator.modules() if hasattr(m, 'Ip')]
    for i, m in enumerate(sn_layers):
        m.u = torch.randn(1, m.weight.size(0), generator=g)
        out['u%d' % i] = m.u.numpy().copy()
    for tag, net in (('g_', env.generator), ('d_', env.discriminator)):
        names = []
        for name, v in net.state_dict().items():
            names.append(name)
            out[tag + 'sd_' + name] = v.numpy().copy()
        out[tag + 'sd_names'] = np.array(names)
    clip = torch.rand(cfg['B'], cfg['K'] + cfg['T'] + cfg['F_'], cfg['c_dim'], cfg['H'], cfg['W'], generator=g) * 2 - 1
    out['clip'] = clip.numpy()
    K, T = cfg['K'], cfg['T']
    env.K, env.T, env.F = cfg['K'], cfg['T'], cfg['F_']
    env.train()
    env.set_train_inputs(clip[:, :K], clip[:, K + T:], clip[:, K:K + T])
    env.forward_train()
    env.optimize_parameters()
    for k in ('Lp', 'gdl', 'L_GAN', 'Lp_forward', 'Lp_backward', 'gdl_forward', 'gdl_backward', 'loss_G', 'loss_d_fake',
              'loss_d_real', 'loss_D'):
        out['loss_' + k] = np.float64(float(getattr(env, k).detach().reshape(-1)[0]))
        print(k, out['loss_' + k])
    gp, dp = dict(env.generator.named_parameters()), dict(env.discriminator.named_parameters())
    gnames = list(gp)
    picks = [gnames[0], gnames[len(gnames) // 2], gnames[-1]]
    for n in picks:
        out['g_after_' + n] = gp[n].detach().numpy().copy()
    out['g_after_names'] = np.array(picks)
    dn = list(dp)
    for n in (dn[0], dn[-2]):
        out['d_after_' + n] = dp[n].detach().numpy().copy()
    out['d_after_names'] = np.array([dn[0], dn[-2]])
    path = os.path.join(HERE, 'tai_step_ref.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path) // 1024, 'KB')


if __name__ == '__main__':
    main()
