"""Shared helpers for the parity tests: seeded inputs and tolerances."""
import numpy as np

from oracle import oracle as O

# FP32 results vs the float64 oracle: |x - ref| <= TOL * max(|ref|, rms(ref))   (north_star: 1e-4 relative)
TOL = 1e-4


def sepconv_inputs(B, C, Ho, Wo, ks, seed=0, padded=True):
    """I ~ U(-1,1); V,H ~ U(-1,1)/sqrt(ks) (signed, unnormalised like the real heads); gO ~ U(-1,1)."""
    rng = np.random.default_rng(seed)
    Hi, Wi = (Ho + ks - 1, Wo + ks - 1) if padded else (Ho, Wo)
    inp = rng.uniform(-1, 1, (B, C, Hi, Wi)).astype(np.float32)
    ver = (rng.uniform(-1, 1, (B, ks, Ho, Wo)) / np.sqrt(ks)).astype(np.float32)
    hor = (rng.uniform(-1, 1, (B, ks, Ho, Wo)) / np.sqrt(ks)).astype(np.float32)
    gout = rng.uniform(-1, 1, (B, C, Ho, Wo)).astype(np.float32)
    return inp, ver, hor, gout


def assert_close(x, ref, tol=TOL, what=""):
    err = O.rel_err(np.asarray(x), np.asarray(ref))
    assert err <= tol, "%s: rel err %.3e > %.1e" % (what, err, tol)
    return err


def to_cuda(*arrays):
    import torch
    return [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in arrays]


def name_seeded_state_dict(template):
    """Deterministic weights for networks too large to commit (tests/golden/make_full_config_golden.py): every
    tensor of `template` (a state_dict; only keys, shapes and dtypes are used) is drawn from a generator seeded
    with crc32(key).  Convolution / linear weights: xavier-normal (gain 1, the reference's weights_init,
    util.py:193-202); biases: uniform(-0.1, 0.1) (the reference zeroes them; non-zero biases keep an O(1) signal in
    every branch of a randomly initialised network)."""
    import math
    import zlib

    import torch
    out = {}
    for key, ref in template.items():
        g = torch.Generator().manual_seed(zlib.crc32(key.encode()))
        shape = tuple(ref.shape)
        if len(shape) >= 2:
            rf = 1
            for s in shape[2:]:
                rf *= s
            std = math.sqrt(2.0 / ((shape[0] + shape[1]) * rf))
            t = torch.randn(shape, generator=g) * std
        else:
            t = torch.rand(shape, generator=g) * 0.2 - 0.1
        out[key] = t.to(ref.dtype)
    return out
