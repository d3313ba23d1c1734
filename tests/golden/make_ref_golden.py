"""Generates tests/golden/sepconv_ref_b200.npz by running the reference's OWN CUDA kernels
(oracle/_ref/libsepconv_ref.so = SeparableConvolution_kernel.cu + SeparableConvolution_cuda.c compiled
unmodified, see oracle/Makefile) on a B200:

    gpurun -- 'python tests/golden/make_ref_golden.py gpurun_out/sepconv_ref_b200.npz'

Inputs are not stored: they are regenerated from the seed by tests.helpers.sepconv_inputs.  The fixture
pins the CPU oracle (tests/test_oracle_cpu.py) to outputs of the reference itself."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import ref_kernels  # noqa: E402
from tests.helpers import sepconv_inputs  # noqa: E402

CASES = [(1, 1, 8, 8, 51), (1, 1, 8, 16, 13), (2, 3, 6, 10, 5), (1, 3, 9, 12, 25), (2, 1, 4, 5, 3)]
SEED = 2024


def main(out_path):
    data = {"cases": np.array(CASES, dtype=np.int64), "seed": np.array(SEED)}
    for n, (B, C, Ho, Wo, ks) in enumerate(CASES):
        inp, ver, hor, gout = sepconv_inputs(B, C, Ho, Wo, ks, seed=SEED + n)
        ti, tv, th, tg = [torch.from_numpy(a).cuda() for a in (inp, ver, hor, gout)]
        out = ref_kernels.forward(ti, tv, th, ks)
        gi, gv, gh = ref_kernels.backward(tg, ti, tv, th, ks)
        torch.cuda.synchronize()
        data["out_%d" % n] = out.cpu().numpy()
        data["gi_%d" % n] = gi.cpu().numpy()
        data["gv_%d" % n] = gv.cpu().numpy()
        data["gh_%d" % n] = gh.cpu().numpy()
    np.savez_compressed(out_path, **data)
    print("wrote", out_path, os.path.getsize(out_path), "bytes;", torch.cuda.get_device_name(0))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "sepconv_ref_b200.npz"))
