"""GPU parity of the decoder resampling kernels (bilinear x2 upsample with the torch-0.3.1 mapping, zero-insertion
unpool + residual add) against the float64 oracle and against the reference's own op spelling on torch CPU.
Integer parts (tap indices, the 2x2 cell positions) are bit-exact; values within 1e-4 relative."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import assert_close, to_cuda

pytestmark = pytest.mark.gpu

UP_SHAPES = [(2, 3, 5, 7), (1, 2, 1, 1), (1, 64, 16, 16), (2, 5, 3, 8), (1, 1, 33, 2), (4, 64, 64, 64),
             # adjoint kernel: W = 4 (both threads of a row use the shifted border window), odd H with several
             # row blocks, a plane count that leaves the last CTA ragged, the UCF decoder shape
             (3, 7, 9, 4), (1, 3, 37, 6), (2, 33, 21, 10), (1, 16, 120, 160), (2, 3, 1, 4), (1, 5, 2, 12)]


@pytest.mark.parametrize("B,C,H,W", UP_SHAPES)
def test_upsample_forward_backward(cuda, B, C, H, W):
    import torch
    import torch.nn.functional as F
    from video_frame_inpainting_b200 import ops
    rng = np.random.default_rng(30)
    x = rng.normal(size=(B, C, H, W)).astype(np.float32)
    g = rng.normal(size=(B, C, 2 * H, 2 * W)).astype(np.float32)
    tx, tg = to_cuda(x, g)
    out = ops.upsample_bilinear2x_forward(tx).cpu().numpy()
    assert out.shape == (B, C, 2 * H, 2 * W)
    assert_close(out, O.upsample_bilinear2x(x), what="upsample fwd")
    # the library op the reference called, evaluated on the CPU with today's spelling of the 0.3.1 mapping
    ref_t = F.interpolate(torch.from_numpy(x), scale_factor=2, mode="bilinear", align_corners=True).numpy()
    assert_close(out, ref_t, what="upsample fwd vs F.interpolate(align_corners=True)")
    gin = ops.upsample_bilinear2x_backward(tg).cpu().numpy()
    assert_close(gin, O.upsample_bilinear2x_backward(g), what="upsample bwd")


def test_upsample_tap_selection_is_bit_exact(cuda):
    """One-hot input rows / columns expose which taps the kernel selected: w0 lands on i0, w1 on i1."""
    from video_frame_inpainting_b200 import ops
    for n in (1, 2, 7, 16, 33):
        i0, i1, w0, w1 = O.upsample_bilinear2x_taps(n)
        eye = np.eye(n, dtype=np.float32).reshape(1, n, n, 1)            # plane c has a 1 in row c, W = 1
        (t,) = to_cuda(eye)
        out = ops.upsample_bilinear2x_forward(t).cpu().numpy()[0, :, :, 0]   # [c, d] = weight of row c in output d
        sel = np.zeros((n, 2 * n))
        np.add.at(sel, (i0, np.arange(2 * n)), w0)
        np.add.at(sel, (i1, np.arange(2 * n)), w1)
        assert np.array_equal(out != 0, sel != 0), "tap index mismatch for n=%d" % n
        assert np.allclose(out, sel, atol=1e-6)


def test_upsample_autograd_and_cpu_guard(cuda):
    import torch
    from video_frame_inpainting_b200 import ops
    x = torch.randn(2, 3, 6, 10, device=cuda, requires_grad=True)
    y = ops.UpsampleBilinear2xFunction.apply(x)
    g = torch.randn_like(y)
    y.backward(g)
    assert_close(x.grad.cpu().numpy(), O.upsample_bilinear2x_backward(g.cpu().numpy()), what="autograd upsample")
    with pytest.raises(NotImplementedError):
        ops.upsample_bilinear2x_forward(torch.zeros(1, 1, 2, 2))


@pytest.mark.parametrize("B,C,H,W", [(2, 3, 5, 7), (1, 1, 1, 1), (2, 128, 32, 32), (1, 4, 9, 6)])
def test_unpool_add_forward_backward(cuda, B, C, H, W):
    import torch
    from video_frame_inpainting_b200 import ops
    rng = np.random.default_rng(31)
    x = rng.normal(size=(B, C, H, W)).astype(np.float32)
    res = rng.normal(size=(B, C, 2 * H, 2 * W)).astype(np.float32)
    g = rng.normal(size=(B, C, 2 * H, 2 * W)).astype(np.float32)
    tx, tr, tg = to_cuda(x, res, g)
    out = ops.unpool_add_forward(tx, tr).cpu().numpy()
    ref = (O.fixed_unpooling(x) + res).astype(np.float32)
    assert np.array_equal(out, ref), "unpool + add is exact in FP32 (one add per element)"
    gx = ops.unpool_backward(tg).cpu().numpy()
    assert np.array_equal(gx, g[:, :, ::2, ::2])
    # autograd: residual gradient is the upstream gradient itself
    a = tx.clone().requires_grad_()
    b = tr.clone().requires_grad_()
    ops.UnpoolAddFunction.apply(a, b).backward(tg)
    assert torch.equal(b.grad, tg) and np.array_equal(a.grad.cpu().numpy(), gx)


@pytest.mark.parametrize("B,C,H,W", [(2, 3, 8, 12), (1, 2, 7, 9), (1, 1, 2, 2), (4, 64, 128, 128), (2, 5, 6, 10), (1, 3, 33, 64)])
def test_maxpool2x2_forward_backward(cuda, B, C, H, W):
    """Values, the 2-bit position codes (bit-exact, ties and NaN included) and the routed gradient; also against
    the library op the reference called (indices of F.max_pool2d on the same device)."""
    import torch
    import torch.nn.functional as F
    from video_frame_inpainting_b200 import ops
    rng = np.random.default_rng(33)
    x = np.maximum(rng.normal(size=(B, C, H, W)), 0).astype(np.float32)   # ReLU output: many exact ties at 0
    x[0, 0, 1, 0] = np.nan
    g = rng.normal(size=(B, C, H // 2, W // 2)).astype(np.float32)
    tx, tg = to_cuda(x, g)
    out, code = ops.maxpool2x2_forward(tx)
    m, c = O.maxpool2x2(x)
    assert np.array_equal(out.cpu().numpy(), m, equal_nan=True)
    assert np.array_equal(code.cpu().numpy(), c)
    gin = ops.maxpool2x2_backward(tg, code, H, W).cpu().numpy()
    assert np.array_equal(gin, O.maxpool2x2_backward(g, c, H, W))
    ref, idx = F.max_pool2d(tx, 2, return_indices=True)
    yy, xx = torch.meshgrid(torch.arange(H // 2, device=cuda), torch.arange(W // 2, device=cuda), indexing="ij")
    flat = (2 * yy + (code // 2).long()) * W + 2 * xx + (code % 2).long()
    assert torch.equal(flat.expand_as(idx), idx)
    # autograd route
    a = torch.from_numpy(np.nan_to_num(x)).cuda().requires_grad_()
    ops.MaxPool2x2Function.apply(a).backward(tg)
    b = torch.from_numpy(np.nan_to_num(x)).cuda().requires_grad_()
    F.max_pool2d(b, 2).backward(tg)
    assert torch.equal(a.grad, b.grad)
