"""Generates tests/golden/frames_u8_ref.npz with the reference's OWN ``save_video_frames`` (predict.py:113-134):
the function is taken verbatim out of /root/reference/predict.py (the module itself cannot be imported: Python-2
imports of datasets / options), executed with ``xrange`` -> ``range`` and the reference's own ``makedir`` /
``to_numpy`` / ``inverse_transform`` (src/util/util.py), and the PNG files it writes are read back.

    python tests/golden/make_frames_golden.py        (build container only)
"""
import ast
import os
import sys
import tempfile

import numpy as np
import torch
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def main():
    from make_model_golden import install_shims
    install_shims()
    sys.path.insert(0, '/root/reference')
    import src.util.util as ref_util
    src = open('/root/reference/predict.py').read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == 'save_video_frames')
    ns = dict(os=os, np=np, torch=torch, Image=Image, xrange=range, makedir=ref_util.makedir,
              to_numpy=ref_util.to_numpy, inverse_transform=ref_util.inverse_transform)
    exec(compile(ast.Module(body=[fn], type_ignores=[]), 'predict.py', 'exec'), ns)
    out = {}
    g = torch.Generator().manual_seed(77)
    for C in (1, 3):
        video = torch.rand(4, C, 16, 24, generator=g) * 2.4 - 1.2          # some values outside [-1, 1]: clamp
        # values on and next to the 8-bit bin edges: x = 2k/255 - 1 (+- one ulp)
        k = torch.arange(0, 256, dtype=torch.float32)
        edges = (2 * k / 255 - 1)
        video[0, 0].view(-1)[:256] = edges
        video[1, 0].view(-1)[:256] = torch.nextafter(edges, torch.full_like(edges, 2.0))
        video[2, 0].view(-1)[:256] = torch.nextafter(edges, torch.full_like(edges, -2.0))
        with tempfile.TemporaryDirectory() as d:
            ns['save_video_frames'](video, d, 'pred_middle', counter_start=5)
            names = sorted(os.listdir(d))
            assert names == ['pred_middle_%04d.png' % (5 + t) for t in range(4)], names
            imgs = np.stack([np.array(Image.open(os.path.join(d, n))) for n in names])
        out['video_c%d' % C] = video.numpy()
        out['png_c%d' % C] = imgs                                            # [T,H,W] (gray) or [T,H,W,3] (RGB)
    np.savez_compressed(os.path.join(HERE, 'frames_u8_ref.npz'), **out)
    print('wrote frames_u8_ref.npz', {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
