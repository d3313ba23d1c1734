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


FULL = [
    (32, 1, 128, 128, 51),   # BASELINE config B operator shape (KTH batch): several tiles per persistent CTA
    (8, 3, 240, 320, 51),    # BASELINE config C operator shape (UCF-101 test frames): chunk-ring forward kernel
    (16, 3, 256, 256, 25),   # op-sweep shape (config E)
    (6, 1, 100, 72, 37),     # ragged: neither dimension a multiple of the tile, ks = 37
]


@pytest.mark.parametrize("B,C,Ho,Wo,ks", FULL)
def test_full_size_against_reference_kernels(cuda, ref, B, C, Ho, Wo, ks):
    """BASELINE.json's full operator shapes: every persistent CTA walks several tiles (halo row ring, chunk ring,
    rolling gI window), which the small oracle-sized cases above do not reach.  Checker = the reference's own
    kernels on the same GPU, compared on the device."""
    import torch
    from video_frame_inpainting_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(77)
    U = lambda *s: torch.rand(*s, device="cuda", generator=g) * 2 - 1
    inp = U(B, C, Ho + ks - 1, Wo + ks - 1)
    ver, hor = U(B, ks, Ho, Wo) / ks ** 0.5, U(B, ks, Ho, Wo) / ks ** 0.5
    gout = U(B, C, Ho, Wo)

    def rel(x, r):
        scale = torch.maximum(r.abs(), r.pow(2).mean().sqrt())
        return float(((x - r).abs() / scale).max())

    r_out = ref.forward(inp, ver, hor, ks)
    r_gi, r_gv, r_gh = ref.backward(gout, inp, ver, hor, ks)
    out = ops.sepconv_forward(inp, ver, hor, ks)
    gi, gv, gh = ops.sepconv_backward(gout, inp, ver, hor, ks)
    for name, x, r in (("fwd", out, r_out), ("gI", gi, r_gi), ("gV", gv, r_gv), ("gH", gh, r_gh)):
        assert rel(x, r) < 2e-4, "%s: rel err %.3e" % (name, rel(x, r))
    # fused pad + 2 x sepconv + blend against the reference op sequence (pad kernel, two launches, blend)
    if ks % 2 == 1:
        pf, pb = U(B, C, Ho, Wo), U(B, C, Ho, Wo)
        ver2, hor2 = U(B, ks, Ho, Wo) / ks ** 0.5, U(B, ks, Ho, Wo) / ks ** 0.5
        pad = torch.nn.ReplicationPad2d(ks // 2)
        d1 = ref.forward(pad(pf).contiguous(), ver, hor, ks)
        d2 = ref.forward(pad(pb).contiguous(), ver2, hor2, ks)
        pred, o1, o2 = ops.tai_fused_forward(pf, pb, ver, hor, ver2, hor2, ks, 0.3, 0.7)
        assert rel(o1, d1) < 2e-4 and rel(o2, d2) < 2e-4
        assert rel(pred, 0.3 * d1 + 0.7 * d2) < 2e-4
