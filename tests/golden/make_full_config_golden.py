"""Generates tests/golden/tai_full_config_ref.npz: the reference's own ``TAIFillInModel(64, 1, 3, 51, num_block=5)``
-- the network registered as ``TAI_gray`` (/root/reference/src/models/create_model.py:27-28: gf_dim 64, ks 51,
five kernel-network blocks) -- run UNMODIFIED on one 64x64 clip with K = F = T = 5, outputs and a few gradients.

Run in the build container only (the reference is not present on the GPU box):

    python tests/golden/make_full_config_golden.py

Same import shims as make_model_golden.py (listed in its header).  The network has 38.3 M parameters (153 MB), far
too many to commit, so the weights are a deterministic function of the state_dict KEY NAMES
(``tests/helpers.py:name_seeded_state_dict``: a generator seeded with crc32(key) per tensor, xavier-normal weights,
uniform(-0.1, 0.1) biases so that every branch of the network carries an O(1) signal); the test re-creates them from
the keys of the product model (whose keys are checked against ``sd_names`` stored here).  The fixture therefore holds
inputs, the five output tensors, the key list and slices of four parameter gradients: a few hundred KB.

Why this fixture exists: the toy networks of tai_model_ref.npz (gf_dim 2-4, ks 5/13) produce outputs of O(1e-6),
which forced loose model-level tolerances.  This one is well conditioned (prediction rms is printed below) and
is checked at 5e-4 (GPU model) / 1e-5 (CPU port) in tests/test_models_{gpu,cpu}.py.
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from make_model_golden import install_shims  # noqa: E402
from tests.helpers import name_seeded_state_dict  # noqa: E402


def main():
    install_shims()
    sys.path.insert(0, '/root/reference')
    from oracle.reference_model import CpuSeparableConvolution
    import src.separable_convolution.SeparableConvolution as ref_op
    ref_op.SeparableConvolution = CpuSeparableConvolution
    import src.models.tai.tai as ref_tai

    torch.set_num_threads(os.cpu_count() or 1)
    cfg = dict(gf_dim=64, c_dim=1, feature_size=3, ks=51, num_block=5, K=5, T=5, F_=5, H=64, W=64, B=1)
    model = ref_tai.TAIFillInModel(cfg['gf_dim'], cfg['c_dim'], cfg['feature_size'], cfg['ks'], num_block=cfg['num_block'])
    model.load_state_dict(name_seeded_state_dict(model.state_dict()), strict=True)
    g = torch.Generator().manual_seed(7)
    # smooth frames (an upsampled noise grid plus a little white noise), in [-1, 1] like the dataset's frames
    def frames(n):
        low = torch.rand(cfg['B'] * n, cfg['c_dim'], 8, 8, generator=g) * 2 - 1
        x = torch.nn.functional.interpolate(low, size=(cfg['H'], cfg['W']), mode='bilinear', align_corners=True)
        x = (0.9 * x + 0.1 * (torch.rand(x.shape, generator=g) * 2 - 1)).clamp(-1, 1)
        return x.view(cfg['B'], n, cfg['c_dim'], cfg['H'], cfg['W']).contiguous()
    pre, fol = frames(cfg['K']), frames(cfg['F_'])
    t0 = time.time()
    res = model(cfg['T'], pre, fol)
    loss = res['pred'].pow(2).mean() + res['pred_forward'].mean() + res['pred_backward'].pow(2).mean()
    loss.backward()
    print('forward + backward: %.1f s' % (time.time() - t0))

    out = {}
    for k, v in cfg.items():
        out['cfg_' + k] = np.int64(v)
    out['pre'], out['fol'] = pre.numpy(), fol.numpy()
    for k, v in res.items():
        out['out_' + k] = v.detach().numpy()
        print('  %-24s rms %.4g  max %.4g' % (k, float(v.detach().pow(2).mean().sqrt()), float(v.detach().abs().max())))
    out['sd_names'] = np.array(list(model.state_dict().keys()))
    names = [n for n, _ in model.named_parameters()]
    picks = [names[0], names[len(names) // 3], names[2 * len(names) // 3], names[-1]]
    params = dict(model.named_parameters())
    for n in picks:
        gr = params[n].grad.numpy()
        out['grad_' + n] = gr.reshape(gr.shape[0], -1)[:8].copy()      # first 8 output channels / rows
        print('  grad %-40s rms %.4g' % (n, float(np.sqrt((gr ** 2).mean()))))
    out['grad_names'] = np.array(picks)
    path = os.path.join(HERE, 'tai_full_config_ref.npz')
    np.savez_compressed(path, **out)
    print('params', sum(p.numel() for p in model.parameters()), 'wrote', path, os.path.getsize(path) // 1024, 'KB')


if __name__ == '__main__':
    main()
