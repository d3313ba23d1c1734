"""Pins the oracle AND the new kernels to the reference's own CUDA kernels (run unmodified on this
GPU from oracle/_ref): the strongest parity evidence available, because the reference ships no
tests or golden vectors for this operator."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import ref_kernels
from tests.helpers import assert_close, sepconv_inputs, to_cuda

pytestmark = pytest.mark.gpu

SHAPES = [(1, 1, 16, 32, 51), (2, 3, 20, 36, 13), (2, 1, 24, 40, 51), (1, 3, 12, 20, 37), (1, 1, 5, 7, 5)]


@pytest.fixture(scope="module")
def ref():
    if not ref_kernels.available():
        pytest.skip("oracle/_ref/libsepconv_ref.so not built (needs /root/reference at build time)")
    return ref_kernels


@pytest.mark.parametrize("B,C,Ho,Wo,ks", SHAPES)
def test_oracle_matches_reference_kernels(cuda, ref, B, C, Ho, Wo, ks):
    inp, ver, hor, gout = sepconv_inputs(B, C, Ho, Wo, ks, seed=11)
    ti, tv, th, tg = to_cuda(inp, ver, hor, gout)
    out = ref.forward(ti, tv, th, ks).cpu().numpy()
    gi, gv, gh = [t.cpu().numpy() for t in ref.backward(tg, ti, tv, th, ks)]
    # FP32 sequential sums of up to 3*51*51 terms vs float64: same 1e-4 bar as the product kernels
    assert_close(out, O.sepconv_forward(inp, ver, hor, ks), what="ref fwd vs oracle")
    assert_close(gv, O.sepconv_grad_vertical(gout, inp, hor, ks), what="ref gV vs oracle")
    assert_close(gh, O.sepconv_grad_horizontal(gout, inp, ver, ks), what="ref gH vs oracle")
    assert_close(gi, O.sepconv_grad_input(gout, ver, hor, ks), what="ref gI vs oracle")


@pytest.mark.parametrize("B,C,Ho,Wo,ks", SHAPES)
def test_new_kernels_match_reference_kernels(cuda, ref, B, C, Ho, Wo, ks):
    from video_frame_inpainting_b200 import ops
    inp, ver, hor, gout = sepconv_inputs(B, C, Ho, Wo, ks, seed=12)
    ti, tv, th, tg = to_cuda(inp, ver, hor, gout)
    r_out = ref.forward(ti, tv, th, ks).cpu().numpy()
    r_gi, r_gv, r_gh = [t.cpu().numpy() for t in ref.backward(tg, ti, tv, th, ks)]
    out = ops.sepconv_forward(ti, tv, th, ks).cpu().numpy()
    gi, gv, gh = [t.cpu().numpy() for t in ops.sepconv_backward(tg, ti, tv, th, ks)]
    for name, x, r in (("fwd", out, r_out), ("gI", gi, r_gi), ("gV", gv, r_gv), ("gH", gh, r_gh)):
        assert_close(x, r, tol=2e-4, what="new vs reference kernel: " + name)  # both FP32, different sum order


def test_reference_tap_count_table(cuda, ref):
    """Integer logic of kernel.cu:150 from the reference itself: ones in -> tap counts out."""
    import torch
    Ho, Wo, ks = 16, 32, 51
    ones = lambda *s: torch.ones(*s, device="cuda")
    gi, _, _ = ref.backward(ones(1, 1, Ho, Wo), ones(1, 1, Ho + ks - 1, Wo + ks - 1), ones(1, ks, Ho, Wo),
                            ones(1, ks, Ho, Wo), ks)
    table = O.sepconv_grad_input_tapcount(Ho + ks - 1, Wo + ks - 1, ks)
    assert np.array_equal(gi[0, 0].cpu().numpy().astype(np.int64), table.astype(np.int64))
