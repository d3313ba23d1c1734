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
