"""GPU parity of the separable-convolution path against the CPU oracle, through the C ABI
(video_frame_inpainting_b200.ops -> libtai_b200.so).  Tolerance: 1e-4 relative (helpers.TOL) for
FP32 values and gradients, bit-exact for everything integer (indices, clamps, bounds)."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import TOL, assert_close, sepconv_inputs, to_cuda

pytestmark = pytest.mark.gpu

# (B, C, Ho, Wo, ks): tiled fast path (ks>=8, Ho>=8, C in {1,3}) and the shape-agnostic fallbacks
SHAPES = [
    (1, 1, 16, 32, 51),   # one tile row, exact tile width
    (2, 1, 24, 40, 51),   # ragged: Wo not a multiple of 32, Ho not a multiple of 8
    (1, 3, 16, 48, 51),   # RGB
    (2, 3, 20, 36, 13),   # J=4 instance
    (1, 1, 9, 33, 25),    # J=7, barely above the tile minimum
    (1, 3, 12, 20, 37),   # J=10, frame narrower than one tile
    (1, 2, 16, 32, 51),   # C=2 -> forward in two channel passes, backward fallback
    (1, 1, 5, 7, 5),      # tiny frame + ks<8 -> fallback kernels
    (1, 1, 8, 8, 63),     # largest tiled ks (J=16)
    (1, 3, 3, 4, 1),      # ks = 1
]


@pytest.mark.parametrize("B,C,Ho,Wo,ks", SHAPES)
def test_forward_matches_oracle(cuda, B, C, Ho, Wo, ks):
    from video_frame_inpainting_b200 import ops
    inp, ver, hor, _ = sepconv_inputs(B, C, Ho, Wo, ks, seed=1)
    ref = O.sepconv_forward(inp, ver, hor, ks)
    ti, tv, th = to_cuda(inp, ver, hor)
    out = ops.sepconv_forward(ti, tv, th, ks).cpu().numpy()
    assert out.shape == ref.shape
    assert_close(out, ref, what="forward %s" % ((B, C, Ho, Wo, ks),))


@pytest.mark.parametrize("B,C,Ho,Wo,ks", SHAPES)
def test_backward_matches_oracle(cuda, B, C, Ho, Wo, ks):
    from video_frame_inpainting_b200 import ops
    inp, ver, hor, gout = sepconv_inputs(B, C, Ho, Wo, ks, seed=2)
    ti, tv, th, tg = to_cuda(inp, ver, hor, gout)
    gi, gv, gh = ops.sepconv_backward(tg, ti, tv, th, ks)
    assert_close(gv.cpu().numpy(), O.sepconv_grad_vertical(gout, inp, hor, ks), what="gradV")
    assert_close(gh.cpu().numpy(), O.sepconv_grad_horizontal(gout, inp, ver, ks), what="gradH")
    assert_close(gi.cpu().numpy(), O.sepconv_grad_input(gout, ver, hor, ks), what="gradI")


def test_autograd_function_contract(cuda):
    """Same call, same arity, gradients flow to all three tensor inputs, None for ks
    (SeparableConvolution.py:11,55,89); CPU tensors raise NotImplementedError (:48-49)."""
    import torch
    from video_frame_inpainting_b200.separable_convolution import SeparableConvolution
    inp, ver, hor, gout = sepconv_inputs(1, 1, 16, 32, 51, seed=3)
    ti, tv, th, tg = to_cuda(inp, ver, hor, gout)
    for t in (ti, tv, th):
        t.requires_grad_(True)
    out = SeparableConvolution.apply(ti, tv, th, 51)
    out.backward(tg)
    assert_close(ti.grad.cpu().numpy(), O.sepconv_grad_input(gout, ver, hor, 51), what="autograd gI")
    assert_close(tv.grad.cpu().numpy(), O.sepconv_grad_vertical(gout, inp, hor, 51), what="autograd gV")
    assert_close(th.grad.cpu().numpy(), O.sepconv_grad_horizontal(gout, inp, ver, 51), what="autograd gH")
    with pytest.raises(NotImplementedError):
        SeparableConvolution.apply(torch.from_numpy(inp), torch.from_numpy(ver), torch.from_numpy(hor), 51)
    with pytest.raises(AssertionError):  # shape algebra of SeparableConvolution.py:27-29
        SeparableConvolution.apply(ti.detach(), tv.detach(), th.detach(), 49)


@pytest.mark.parametrize("ks,C", [(51, 1), (13, 3), (5, 2)])
def test_one_hot_kernels_select_input_pixel_bit_exact(cuda, ks, C):
    """KAT: V = e_i0, H = e_j0  =>  O[y,x] == I[y+i0, x+j0] exactly (index logic is integer)."""
    from video_frame_inpainting_b200 import ops
    B, Ho, Wo = 2, 16, 40
    inp, _, _, _ = sepconv_inputs(B, C, Ho, Wo, ks, seed=4)
    rng = np.random.default_rng(5)
    i0 = rng.integers(0, ks, (B, Ho, Wo))
    j0 = rng.integers(0, ks, (B, Ho, Wo))
    ver = np.zeros((B, ks, Ho, Wo), np.float32)
    hor = np.zeros((B, ks, Ho, Wo), np.float32)
    bb, yy, xx = np.meshgrid(np.arange(B), np.arange(Ho), np.arange(Wo), indexing="ij")
    ver[bb, i0, yy, xx] = 1
    hor[bb, j0, yy, xx] = 1
    expect = inp[bb[:, None], np.arange(C)[None, :, None, None], (yy + i0)[:, None], (xx + j0)[:, None]]
    ti, tv, th = to_cuda(inp, ver, hor)
    out = ops.sepconv_forward(ti, tv, th, ks).cpu().numpy()
    assert np.array_equal(out, expect)


@pytest.mark.parametrize("Ho,Wo,ks", [(16, 32, 51), (24, 40, 13), (5, 7, 5)])
def test_grad_input_tap_count_bit_exact(cuda, Ho, Wo, ks):
    """With gO = V = H = 1 the grad-input kernel counts the taps that pass the bounds test of
    kernel.cu:150: an integer table, compared bit-exactly with the oracle's."""
    import torch
    from video_frame_inpainting_b200 import ops
    B, C = 1, 1
    ones = lambda *s: torch.ones(*s, device="cuda")
    gi, _, _ = ops.sepconv_backward(ones(B, C, Ho, Wo), ones(B, C, Ho + ks - 1, Wo + ks - 1),
                                    ones(B, ks, Ho, Wo), ones(B, ks, Ho, Wo), ks, needs=(True, False, False))
    table = O.sepconv_grad_input_tapcount(Ho + ks - 1, Wo + ks - 1, ks)
    assert np.array_equal(gi[0, 0].cpu().numpy().astype(np.int64), table.astype(np.int64))
    assert table.sum() == Ho * Wo * ks * ks  # every tap of every output pixel lands exactly once


@pytest.mark.parametrize("H,W,p", [(16, 32, 25), (5, 7, 2), (1, 1, 3), (9, 4, 0), (2, 6, 25), (3, 300, 4), (1, 40, 7)])
def test_replication_pad_index_bit_exact(cuda, H, W, p):
    """Clamp logic of ReplicationPad2d (tai.py:170-171): padding an index image must reproduce the
    oracle's integer tables exactly; the adjoint of a ones image must give the integer multiplicity."""
    import torch
    from video_frame_inpainting_b200 import ops
    idx = np.arange(H * W, dtype=np.float32).reshape(1, 1, H, W)
    padded = ops.replication_pad_forward(torch.from_numpy(idx).cuda(), p).cpu().numpy()
    sy, sx = O.replication_pad_index(H, W, p)
    assert np.array_equal(padded[0, 0].astype(np.int64), (sy[:, None].astype(np.int64) * W + sx[None, :]))
    mult = ops.replication_pad_backward(torch.ones(1, 1, H + 2 * p, W + 2 * p, device="cuda"), p).cpu().numpy()
    ref = O.replication_pad_adjoint(np.ones((1, 1, H + 2 * p, W + 2 * p)), p)
    assert np.array_equal(mult, ref.astype(np.float32))


@pytest.mark.parametrize("N,H,W,p", [(6, 16, 32, 25), (3, 2, 5, 3), (4, 1, 9, 2), (32, 128, 128, 25), (2, 7, 270, 6)])
def test_replication_pad_adjoint_values(cuda, N, H, W, p):
    """Random gradients, several images: border-row CTAs and interior-row warps of the adjoint kernel."""
    from video_frame_inpainting_b200 import ops
    rng = np.random.default_rng(9)
    g = rng.normal(size=(N, 1, H + 2 * p, W + 2 * p)).astype(np.float32)
    (tg,) = to_cuda(g)
    out = ops.replication_pad_backward(tg, p).cpu().numpy()
    assert_close(out, O.replication_pad_adjoint(g, p), what="pad adjoint")
    assert np.array_equal(out, ops.replication_pad_backward(tg, p).cpu().numpy()), "deterministic"


@pytest.mark.parametrize("B,C,H,W,ks,a,b", [
    (1, 1, 16, 32, 51, 0.5, 0.5),      # TAI blend (tai.py:105)
    (2, 3, 24, 40, 51, 0.75, 0.25),    # TWI blend (twi.py:105), w = 0.25
    (1, 3, 16, 36, 13, 1.0 / 3, 2.0 / 3),
    (1, 1, 5, 7, 5, 0.5, 0.5),         # fallback
    (2, 1, 32, 64, 13, 0.5, 0.5),      # small ks: the 7 prologue rows span two TMA tap chunks
    (1, 1, 16, 32, 25, 0.5, 0.5),
    (1, 3, 24, 32, 37, 0.5, 0.5),
])
def test_fused_pad_sepconv_blend(cuda, B, C, H, W, ks, a, b):
    from video_frame_inpainting_b200 import ops
    rng = np.random.default_rng(7)
    pf = rng.uniform(-1, 1, (B, C, H, W)).astype(np.float32)
    pb = rng.uniform(-1, 1, (B, C, H, W)).astype(np.float32)
    maps = [(rng.uniform(-1, 1, (B, ks, H, W)) / np.sqrt(ks)).astype(np.float32) for _ in range(4)]
    ref_pred, ref_d1, ref_d2 = O.tai_fused_forward(pf, pb, *maps, ks, a, b)
    t = to_cuda(pf, pb, *maps)
    pred, d1, d2 = ops.tai_fused_forward(*t, ks, a, b)
    assert_close(d1.cpu().numpy(), ref_d1, what="dot1")
    assert_close(d2.cpu().numpy(), ref_d2, what="dot2")
    assert_close(pred.cpu().numpy(), ref_pred, what="pred")
    # and the unfused route through the reference-shaped operator gives the same thing
    unf = ops.sepconv_forward(ops.replication_pad_forward(t[0], ks // 2), t[2], t[3], ks)
    assert_close(unf.cpu().numpy(), ref_d1, what="pad+op")
    pred_only, n1, n2 = ops.tai_fused_forward(*t, ks, a, b, emit_intermediate=False)
    assert n1 is None and n2 is None
    assert np.array_equal(pred_only.cpu().numpy(), pred.cpu().numpy())


@pytest.mark.parametrize("B,C,H,W,ks", [
    (1, 1, 16, 32, 51), (2, 3, 12, 36, 13), (1, 1, 5, 7, 5), (2, 1, 32, 64, 13),
    # several tile columns (interior tiles + both borders), tiles that touch both borders at once, ragged
    # sizes, three channels, ks = 25 / 37; H = 1 / 2: every row of the pad adjoint is a border row
    (1, 1, 40, 128, 51), (1, 3, 24, 100, 51), (2, 1, 9, 20, 25), (1, 3, 17, 76, 37), (1, 1, 70, 36, 13),
    (2, 1, 1, 12, 5), (1, 2, 2, 9, 13)])
def test_fused_backward(cuda, B, C, H, W, ks):
    import torch
    from video_frame_inpainting_b200 import ops
    a, b = 0.6, 0.4
    p = ks // 2
    rng = np.random.default_rng(8)
    pf = rng.uniform(-1, 1, (B, C, H, W)).astype(np.float32)
    pb = rng.uniform(-1, 1, (B, C, H, W)).astype(np.float32)
    maps = [(rng.uniform(-1, 1, (B, ks, H, W)) / np.sqrt(ks)).astype(np.float32) for _ in range(4)]
    g_pred = rng.uniform(-1, 1, (B, C, H, W)).astype(np.float32)
    g_d1 = rng.uniform(-1, 1, (B, C, H, W)).astype(np.float32)
    t = [x.requires_grad_(True) for x in to_cuda(pf, pb, *maps)]
    pred, d1, d2 = ops.tai_blend_sepconv(*t, ks, a, b)
    tg, tg1 = to_cuda(g_pred, g_d1)
    (pred * tg).sum().add((d1 * tg1).sum()).backward()
    gD1 = (a * g_pred.astype(np.float64) + g_d1).astype(np.float32)
    gD2 = (b * g_pred.astype(np.float64)).astype(np.float32)
    for s, (src, gD, v, h) in enumerate(((pf, gD1, maps[0], maps[1]), (pb, gD2, maps[2], maps[3]))):
        padded = O.replication_pad(src, p).astype(np.float32)
        assert_close(t[2 + 2 * s].grad.cpu().numpy(), O.sepconv_grad_vertical(gD, padded, h, ks), tol=2 * TOL, what="fused gV%d" % s)
        assert_close(t[3 + 2 * s].grad.cpu().numpy(), O.sepconv_grad_horizontal(gD, padded, v, ks), tol=2 * TOL, what="fused gH%d" % s)
        gi = O.replication_pad_adjoint(O.sepconv_grad_input(gD, v, h, ks), p)
        assert_close(t[s].grad.cpu().numpy(), gi, tol=2 * TOL, what="fused gPred%d" % s)


def test_fused_backward_full_size_adjoint_identity(cuda):
    """BASELINE config B size through the fused operator: <g, blend(pf, pb)> == <g_pf, pf> + <g_pb, pb>
    (the operator is linear in the two predictions; checks gI scatter + pad adjoint at 128 x 128, B = 32)."""
    import torch
    from video_frame_inpainting_b200 import ops
    B, C, H, W, ks = 32, 1, 128, 128, 51
    g = torch.Generator(device="cuda").manual_seed(1)
    U = lambda *s: torch.rand(*s, device="cuda", generator=g) * 2 - 1
    pf, pb = U(B, C, H, W).requires_grad_(), U(B, C, H, W).requires_grad_()
    maps = [U(B, ks, H, W) / ks ** 0.5 for _ in range(4)]
    pred, _, _ = ops.tai_blend_sepconv(pf, pb, *maps, ks, 0.5, 0.5)
    go = U(B, C, H, W)
    (pred * go).sum().backward()
    dot = lambda x, y: (x.double() * y.double()).sum().item()
    lhs, rhs = dot(go, pred.detach()), dot(pf.grad, pf.detach()) + dot(pb.grad, pb.detach())
    assert abs(lhs - rhs) <= 1e-5 * max(1.0, abs(lhs)) + 1e-3, (lhs, rhs)
    # one sample against the oracle
    gi = O.replication_pad_adjoint(O.sepconv_grad_input(0.5 * go[3:4].cpu().numpy(), maps[0][3:4].cpu().numpy(),
                                                        maps[1][3:4].cpu().numpy(), ks), ks // 2)
    assert_close(pf.grad[3:4].cpu().numpy(), gi, tol=2 * TOL, what="full-size fused gPred")


def test_full_size_properties_kth_batch(cuda):
    """BASELINE config B size (B=32, C=1, 128x128, ks=51), checked through size-independent
    properties: linearity in I, the adjoint identity <gO, fwd(I)> == <gI, I>, and
    <gV, V> == <gH, H> == <gO, O> (the output is linear in V and in H)."""
    import torch
    from video_frame_inpainting_b200 import ops
    B, C, Ho, Wo, ks = 32, 1, 128, 128, 51
    g = torch.Generator(device="cuda").manual_seed(0)
    U = lambda *s: torch.rand(*s, device="cuda", generator=g) * 2 - 1
    I1, I2 = U(B, C, Ho + ks - 1, Wo + ks - 1), U(B, C, Ho + ks - 1, Wo + ks - 1)
    V, H = U(B, ks, Ho, Wo) / ks ** 0.5, U(B, ks, Ho, Wo) / ks ** 0.5
    gO = U(B, C, Ho, Wo)
    o1, o2 = ops.sepconv_forward(I1, V, H, ks), ops.sepconv_forward(I2, V, H, ks)
    o12 = ops.sepconv_forward(I1 + 2 * I2, V, H, ks)
    scale = o12.double().pow(2).mean().sqrt().item()
    assert (o12 - (o1 + 2 * o2)).abs().max().item() <= 20 * TOL * scale
    gi, gv, gh = ops.sepconv_backward(gO, I1, V, H, ks)
    dot = lambda x, y: (x.double() * y.double()).sum().item()
    lhs = dot(gO, o1)
    for name, val in (("gI", dot(gi, I1)), ("gV", dot(gv, V)), ("gH", dot(gh, H))):
        assert abs(val - lhs) <= 1e-5 * max(1.0, abs(lhs)) + 1e-3, (name, val, lhs)
    # spot-check one sample of the batch against the oracle
    ref = O.sepconv_forward(I1[5:6].cpu().numpy(), V[5:6].cpu().numpy(), H[5:6].cpu().numpy(), ks)
    assert_close(o1[5:6].cpu().numpy(), ref, what="full-size sample")


def test_error_codes_do_not_throw_across_the_abi(cuda):
    import torch
    from video_frame_inpainting_b200 import _lib
    x = torch.zeros(16, device="cuda")
    with pytest.raises(_lib.TaiB200Error) as ei:
        _lib.call("SeparableConvolution_cuda_forward_b200", x.data_ptr(), x.data_ptr(), x.data_ptr(), x.data_ptr(),
                  1, 1, 3, 3, 5, None)
    assert ei.value.code == -1
    with pytest.raises(_lib.TaiB200Error):
        _lib.call("tai_fused_forward_b200", *([x.data_ptr()] * 9), 1, 1, 4, 4, 4, 0.5, 0.5, None)  # even ks
