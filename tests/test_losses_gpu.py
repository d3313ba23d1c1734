"""GPU parity of the fused MSE + GDL loss kernels against the float64 oracle and
against golden vectors produced by the reference's own GDL class (tests/golden/make_loss_golden.py).
Values / gradients within 1e-4 relative; the sign selection of the backward is exact (checked on ties)."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import assert_close, to_cuda

pytestmark = pytest.mark.gpu

SHAPES = [(3, 2, 1, 5, 7), (2, 3, 12, 16), (160, 1, 128, 128), (4, 1, 9, 4), (1, 1, 2, 2), (2, 1, 33, 21), (7, 3, 40, 52),
          (24, 3, 240, 320), (1, 1, 300, 8)]


@pytest.mark.parametrize("shape", SHAPES)
def test_l2_gdl_forward_backward(cuda, shape):
    import torch
    from video_frame_inpainting_b200 import ops
    rng = np.random.default_rng(40)
    x = rng.uniform(-1, 1, shape).astype(np.float32)
    y = rng.uniform(-1, 1, shape).astype(np.float32)
    tx, ty = to_cuda(x, y)
    out = ops.l2_gdl_loss_forward(tx, ty).cpu().numpy()
    mse, gdl = O.l2_gdl_loss(x, y)
    assert abs(out[0] - mse) <= 1e-4 * abs(mse) and abs(out[1] - gdl) <= 1e-4 * abs(gdl), (out, mse, gdl)
    gm = torch.tensor([0.7], device=cuda)
    gg = torch.tensor([1.3], device=cuda)
    g = ops.l2_gdl_loss_backward(tx, ty, gm, gg).cpu().numpy()
    assert_close(g, O.l2_gdl_loss_backward(x, y, 0.7, 1.3), what="l2+gdl grad")
    # determinism: fixed-order final sum
    assert np.array_equal(out, ops.l2_gdl_loss_forward(tx, ty).cpu().numpy())


def test_l2_gdl_matches_reference_golden(cuda):
    import torch
    from video_frame_inpainting_b200 import ops
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "l2_gdl_ref.npz"))
    for i in range(int(z["n"])):
        tx, ty = to_cuda(z["x%d" % i], z["y%d" % i])
        tx.requires_grad_()
        mse, gdl = ops.l2_gdl_loss(tx, ty)
        assert abs(mse.item() - float(z["mse%d" % i])) <= 1e-5 * abs(mse.item())
        assert abs(gdl.item() - float(z["gdl%d" % i])) <= 1e-5 * abs(gdl.item())
        (float(z["g_mse"]) * mse + float(z["g_gdl"]) * gdl).backward()
        assert_close(tx.grad.cpu().numpy(), z["grad%d" % i], what="autograd l2+gdl vs reference GDL")


def test_l2_gdl_sign_ties_and_unaligned(cuda):
    """Constant regions make |.| arguments exactly zero: the gradient there is the MSE part only (sign(0) = 0).
    A view that is not 16 B-aligned takes the scalar kernels."""
    import torch
    from video_frame_inpainting_b200 import ops
    x = np.full((2, 1, 8, 12), 0.25, np.float32)
    y = np.full((2, 1, 8, 12), -0.5, np.float32)
    tx, ty = to_cuda(x, y)
    one, zero = torch.ones(1, device=cuda), torch.zeros(1, device=cuda)
    g = ops.l2_gdl_loss_backward(tx, ty, one, one).cpu().numpy()
    assert np.array_equal(g, ops.l2_gdl_loss_backward(tx, ty, one, zero).cpu().numpy()), "sign(0) must be 0"
    assert_close(g, O.l2_gdl_loss_backward(x, y, 1.0, 1.0), what="tie gradient")
    rng = np.random.default_rng(41)
    big = torch.from_numpy(rng.uniform(-1, 1, 2 * 8 * 12 + 1).astype(np.float32)).cuda()
    vx = big[1:].view(2, 1, 8, 12)  # 4 B-aligned only
    out = ops.l2_gdl_loss_forward(vx, ty).cpu().numpy()
    mse, gdl = O.l2_gdl_loss(vx.cpu().numpy(), y)
    assert abs(out[0] - mse) <= 1e-4 * mse and abs(out[1] - gdl) <= 1e-4 * gdl
    with pytest.raises(NotImplementedError):
        ops.l2_gdl_loss_forward(torch.zeros(1, 1, 4, 4), torch.zeros(1, 1, 4, 4))


def test_fused_loss_equals_reference_composition_on_device(cuda):
    """The route the training environment takes (L2GDLLoss on [B,T,C,H,W]) against the reference's spelling
    on the same device: time-major regrouping, inverse_transform, MSELoss, GDL (environments.py:363-371)."""
    import torch
    from video_frame_inpainting_b200.environments.environments import L2GDLDiscTrainingEnvironment as Env
    from video_frame_inpainting_b200.losses.losses import GDL, L2GDLLoss
    torch.manual_seed(5)
    pred = (torch.rand(3, 4, 1, 32, 48, device=cuda) * 2 - 1).requires_grad_()
    gt = torch.rand(4, 9, 1, 32, 48, device=cuda)[1:, 2:6] * 2 - 1            # a non-contiguous slice, as in the step
    mse, gdl = L2GDLLoss()(pred, gt)
    (mse + gdl).backward()
    g_fused = pred.grad.clone()
    pred.grad = None
    a, b = Env._time_major01(pred), Env._time_major01(gt)
    mse_ref, gdl_ref = torch.nn.MSELoss()(a, b), GDL()(a, b)
    (mse_ref + gdl_ref).backward()
    assert abs(mse.item() - mse_ref.item()) <= 1e-5 * mse_ref.item()
    assert abs(gdl.item() - gdl_ref.item()) <= 1e-5 * gdl_ref.item()
    assert_close(g_fused.cpu().numpy(), pred.grad.cpu().numpy(), what="fused loss gradient vs torch composition")
