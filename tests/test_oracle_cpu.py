"""CPU tests of the oracle (test infrastructure): the C restatement against an independent NumPy
restatement, against known-answer tests, against fixtures produced by the reference's own CUDA kernels on
a B200 (tests/golden/sepconv_ref_b200.npz), and -- for the torch-0.3.1 library ops on the path -- against
the torch 2.x CPU ops whose semantics are unchanged."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import oracle as O
from tests.helpers import sepconv_inputs

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "sepconv_ref_b200.npz")


@pytest.mark.parametrize("B,C,Ho,Wo,ks", [(2, 3, 6, 7, 5), (1, 1, 4, 9, 13), (1, 2, 3, 3, 1), (1, 1, 2, 2, 51)])
def test_c_oracle_matches_numpy_restatement(B, C, Ho, Wo, ks):
    inp, ver, hor, gout = sepconv_inputs(B, C, Ho, Wo, ks, seed=3)
    assert np.allclose(O.sepconv_forward(inp, ver, hor, ks), O.sepconv_forward_np(inp, ver, hor, ks), atol=1e-13)
    gi, gv, gh = O.sepconv_backward_np(gout, inp, ver, hor, ks)
    assert np.allclose(O.sepconv_grad_input(gout, ver, hor, ks), gi, atol=1e-13)
    assert np.allclose(O.sepconv_grad_vertical(gout, inp, hor, ks), gv, atol=1e-13)
    assert np.allclose(O.sepconv_grad_horizontal(gout, inp, ver, ks), gh, atol=1e-13)
    # the FP32 "port" flavour differs from float64 only by rounding
    assert O.rel_err(O.sepconv_forward(inp, ver, hor, ks, np.float32), O.sepconv_forward(inp, ver, hor, ks)) < 1e-5


def test_oracle_matches_reference_kernel_fixtures():
    """The fixture holds outputs of the reference's unmodified kernels (kernel.cu:19-162) run on a B200."""
    z = np.load(GOLDEN)
    seed = int(z["seed"])
    for n, (B, C, Ho, Wo, ks) in enumerate(z["cases"].tolist()):
        inp, ver, hor, gout = sepconv_inputs(B, C, Ho, Wo, ks, seed=seed + n)
        for name, ref in (("out", O.sepconv_forward(inp, ver, hor, ks)),
                          ("gi", O.sepconv_grad_input(gout, ver, hor, ks)),
                          ("gv", O.sepconv_grad_vertical(gout, inp, hor, ks)),
                          ("gh", O.sepconv_grad_horizontal(gout, inp, ver, ks))):
            got = z["%s_%d" % (name, n)]
            assert got.shape == ref.shape
            assert O.rel_err(got, ref) < 1e-4, (name, n)
        # and the FP32 port, which sums in the reference's order, is closer still
        assert O.rel_err(z["out_%d" % n], O.sepconv_forward(inp, ver, hor, ks, np.float32)) < 2e-5


def test_known_answers():
    B, C, Ho, Wo, ks = 1, 2, 5, 6, 7
    inp, _, _, gout = sepconv_inputs(B, C, Ho, Wo, ks, seed=5)
    # one-hot kernels select one input pixel
    ver = np.zeros((B, ks, Ho, Wo), np.float32)
    hor = np.zeros((B, ks, Ho, Wo), np.float32)
    ver[:, 2] = 1
    hor[:, 5] = 1
    assert np.array_equal(O.sepconv_forward(inp, ver, hor, ks), inp[:, :, 2:2 + Ho, 5:5 + Wo].astype(np.float64))
    # constant 1/ks kernels are a box filter
    box = np.full((B, ks, Ho, Wo), 1.0 / ks, np.float32)
    win = np.lib.stride_tricks.sliding_window_view(inp.astype(np.float64), (ks, ks), axis=(2, 3))
    assert np.allclose(O.sepconv_forward(inp, box, box, ks), win.mean(axis=(-1, -2)) * (ks * np.float64(np.float32(1 / ks))) ** 2, atol=1e-12)
    # linearity in I, adjoint identity <gO, fwd(I)> == <gI, I>, and <gV,V> == <gH,H> == <gO,O>
    _, v, h, _ = sepconv_inputs(B, C, Ho, Wo, ks, seed=6)
    o = O.sepconv_forward(inp, v, h, ks)
    assert np.allclose(O.sepconv_forward(2 * inp, v, h, ks), 2 * o)
    lhs = float((gout * o).sum())
    assert np.isclose((O.sepconv_grad_input(gout, v, h, ks) * inp).sum(), lhs)
    assert np.isclose((O.sepconv_grad_vertical(gout, inp, h, ks) * v).sum(), lhs)
    assert np.isclose((O.sepconv_grad_horizontal(gout, inp, v, ks) * h).sum(), lhs)


def test_finite_differences_of_the_kernel_maps():
    B, C, Ho, Wo, ks = 1, 1, 3, 4, 5
    inp, v, h, gout = sepconv_inputs(B, C, Ho, Wo, ks, seed=7)
    gv = O.sepconv_grad_vertical(gout, inp, h, ks)
    gh = O.sepconv_grad_horizontal(gout, inp, v, ks)
    loss = lambda vv, hh: float((O.sepconv_forward_np(inp, vv, hh, ks) * gout).sum())
    eps = 1e-3
    rng = np.random.default_rng(0)
    for _ in range(10):
        idx = tuple(rng.integers(0, s) for s in v.shape)
        dv = v.astype(np.float64).copy(); dv[idx] += eps
        dh = h.astype(np.float64).copy(); dh[idx] += eps
        assert np.isclose((loss(dv, h) - loss(v, h)) / eps, gv[idx], rtol=1e-6, atol=1e-9)
        assert np.isclose((loss(v, dh) - loss(v, h)) / eps, gh[idx], rtol=1e-6, atol=1e-9)


def test_integer_tables():
    cnt = O.sepconv_grad_input_tapcount(10, 12, 5)
    assert cnt.sum() == 6 * 8 * 25 and cnt.max() == 25 and cnt[0, 0] == 1 and cnt[4, 4] == 25
    sy, sx = O.replication_pad_index(4, 3, 2)
    assert sy.tolist() == [0, 0, 0, 1, 2, 3, 3, 3] and sx.tolist() == [0, 0, 0, 1, 2, 2, 2]


def test_replication_pad_and_adjoint_against_torch():
    x = torch.randn(2, 3, 5, 7, dtype=torch.float64, requires_grad=True)
    p = 3
    y = torch.nn.ReplicationPad2d([p, p, p, p])(x)
    assert np.array_equal(O.replication_pad(x.detach().numpy(), p), y.detach().numpy())
    g = torch.randn_like(y)
    y.backward(g)
    assert np.allclose(O.replication_pad_adjoint(g.numpy(), p), x.grad.numpy())


def test_gates_against_torch_chain():
    """mcnet.py:287-293 spelled with torch 2.x ops (sigmoid / tanh semantics are unchanged since 0.3.1)."""
    conv = torch.randn(2, 16, 3, 5, dtype=torch.float64, requires_grad=True)
    state = torch.randn(2, 8, 3, 5, dtype=torch.float64, requires_grad=True)
    c, h = torch.chunk(state, 2, dim=1)
    i, j, f, o = torch.chunk(conv, 4, dim=1)
    new_c = c * torch.sigmoid(f + 1) + torch.sigmoid(i) * torch.tanh(j)
    new_h = torch.tanh(new_c) * torch.sigmoid(o)
    new_state = torch.cat((new_c, new_h), dim=1)
    nh, ns = O.convlstm_gates(conv.detach().numpy(), state.detach().numpy(), 1.0)
    assert np.allclose(ns, new_state.detach().numpy()) and np.allclose(nh, new_h.detach().numpy())
    g = torch.randn_like(new_state)
    new_state.backward(g)
    gc, gs = O.convlstm_gates_backward(conv.detach().numpy(), state.detach().numpy(), g.numpy(), 1.0)
    assert np.allclose(gc, conv.grad.numpy()) and np.allclose(gs, state.grad.numpy())


def _reference_warp(img, uv):
    """slomo.py:270-284 with today's spelling of the 0.3.1 sampler (align_corners=True, zeros)."""
    H, W = img.shape[-2:]
    X = torch.arange(W, dtype=img.dtype)[None, None, :] + uv[:, 0]
    Y = torch.arange(H, dtype=img.dtype)[None, :, None] + uv[:, 1]
    grid = torch.stack((2 * (X / W - 0.5), 2 * (Y / H - 0.5)), dim=3)
    return F.grid_sample(img, grid, mode="bilinear", padding_mode="zeros", align_corners=True)


def test_warp_against_grid_sample():
    rng = np.random.default_rng(1)
    img = rng.uniform(-1, 1, (2, 3, 9, 11))
    uv = rng.normal(0, 3, (2, 2, 9, 11))
    ti, tu = torch.tensor(img, requires_grad=True), torch.tensor(uv, requires_grad=True)
    out = _reference_warp(ti, tu)
    assert np.allclose(O.flow_warp(img, uv), out.detach().numpy(), atol=1e-12)
    assert O.rel_err(O.flow_warp(img.astype(np.float32), uv.astype(np.float32), coords="f32"), out.detach().numpy()) < 1e-4
    g = rng.uniform(-1, 1, out.shape)
    out.backward(torch.tensor(g))
    gi, gu = O.flow_warp_backward(img, uv, g)
    assert np.allclose(gi, ti.grad.numpy(), atol=1e-10) and np.allclose(gu, tu.grad.numpy(), atol=1e-10)
    # a zero flow is NOT an identity warp in the reference (samples x*(W-1)/W)
    ident = O.flow_warp(img, np.zeros_like(uv))
    assert not np.allclose(ident, img)


def test_slomo_stage_formulas():
    rng = np.random.default_rng(2)
    f01, f10 = rng.normal(size=(1, 2, 4, 4)), rng.normal(size=(1, 2, 4, 4))
    T = 3
    for t_ in range(T):
        t = (t_ + 1) / (T + 1)
        a, b = O.slomo_flow_combine(f01, f10, t)
        assert np.allclose(a, -(1 - t) * t * f01 + t * t * f10) and np.allclose(b, (1 - t) ** 2 * f01 - t * (1 - t) * f10)
    assert O.time_weights(3) == [0.25, 0.5, 0.75]
    assert np.allclose(O.tai_blend(np.ones(3), 3 * np.ones(3)), 2.0)


def test_upsample_and_unpool_against_torch():
    """The decoder resampling oracles against the library ops the reference called (torch CPU, today's spelling
    of the 0.3.1 mapping) and against the reference's own cat / permute spelling of fixed_unpooling."""
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(9)
    for shape in [(2, 3, 5, 7), (1, 2, 1, 1), (1, 1, 16, 16), (2, 2, 3, 8)]:
        x = rng.normal(size=shape).astype(np.float32)
        xt = torch.from_numpy(x).double().requires_grad_()
        y = F.interpolate(xt, scale_factor=2, mode='bilinear', align_corners=True)
        assert O.rel_err(O.upsample_bilinear2x(x), y.detach().numpy()) < 2e-5   # FP32 weights vs float64 weights
        g = rng.normal(size=y.shape)
        y.backward(torch.from_numpy(g))
        assert O.rel_err(O.upsample_bilinear2x_backward(g), xt.grad.numpy()) < 2e-5
        # adjoint identity <up(x), g> == <x, up^T(g)>
        assert abs(np.sum(O.upsample_bilinear2x(x) * g) - np.sum(x * O.upsample_bilinear2x_backward(g))) < 1e-9 * g.size
        # mcnet.py:240-256 verbatim
        p = torch.from_numpy(x).permute(0, 2, 3, 1)
        out = torch.cat((p, p.clone().zero_()), dim=3)
        out = torch.cat((out, out.clone().zero_()), dim=2)
        ref = out.view(p.size(0), 2 * p.size(1), 2 * p.size(2), p.size(3)).permute(0, 3, 1, 2).numpy()
        assert np.array_equal(O.fixed_unpooling(x), ref)
    i0, i1, w0, w1 = O.upsample_bilinear2x_taps(16)
    assert i0[0] == 0 and i1[-1] == 15 and i0[-1] == 15 and abs(w0[-1] + w1[-1] - 1) < 1e-7


def test_loss_oracle_matches_reference_gdl_golden():
    """tests/golden/l2_gdl_ref.npz was produced by the reference's own GDL class + torch.nn.MSELoss
    (tests/golden/make_loss_golden.py); the oracle must reproduce values and gradients, ties included."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "l2_gdl_ref.npz"))
    for i in range(int(z["n"])):
        x, y = z["x%d" % i], z["y%d" % i]
        mse, gdl = O.l2_gdl_loss(x, y)
        assert abs(mse - float(z["mse%d" % i])) <= 1e-6 * abs(mse)
        assert abs(gdl - float(z["gdl%d" % i])) <= 1e-6 * abs(gdl)
        g = O.l2_gdl_loss_backward(x, y, float(z["g_mse"]), float(z["g_gdl"]))
        assert O.rel_err(z["grad%d" % i], g) < 1e-5
    # the product's torch GDL module (CPU / API parity with losses.py:4-45) agrees as well
    import torch
    from video_frame_inpainting_b200.losses.losses import GDL
    a = (torch.from_numpy(z["x0"]) + 1.) / 2
    b = (torch.from_numpy(z["y0"]) + 1.) / 2
    assert abs(GDL()(a, b).item() - float(z["gdl0"])) < 1e-6
    assert GDL(reduce=False)(a, b).shape == a.shape[:-2] + (a.shape[-2] - 1, a.shape[-1] - 1)


def test_maxpool_oracle_matches_library_op_including_ties():
    """Tie rule of nn.MaxPool2d (first maximum in scan order, NaN wins): the oracle's codes must reproduce the
    indices torch's CPU kernel returns, on ReLU-like inputs full of equal zeros and on odd sizes."""
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(11)
    for shape in [(2, 3, 8, 12), (1, 2, 7, 9), (1, 1, 2, 2)]:
        x = np.maximum(rng.normal(size=shape), 0).astype(np.float32)     # ~half of the elements are exactly 0
        x[0, 0, 0, 1] = np.nan
        m, code = O.maxpool2x2(x)
        ref, idx = F.max_pool2d(torch.from_numpy(x), 2, return_indices=True)
        assert np.array_equal(m, ref.numpy(), equal_nan=True)
        H, W = shape[-2:]
        yy, xx = np.meshgrid(np.arange(H // 2), np.arange(W // 2), indexing="ij")
        flat = (2 * yy + code // 2) * W + 2 * xx + code % 2
        assert np.array_equal(flat, idx.numpy())
        g = rng.normal(size=m.shape).astype(np.float32)
        t = torch.from_numpy(np.nan_to_num(x)).requires_grad_()
        F.max_pool2d(t, 2).backward(torch.from_numpy(g))
        _, code2 = O.maxpool2x2(np.nan_to_num(x))
        assert np.array_equal(O.maxpool2x2_backward(g, code2, H, W), t.grad.numpy())


def test_frames_to_uint8_oracle_matches_reference_save_video_frames():
    """tests/golden/frames_u8_ref.npz: PNGs written by the reference's own save_video_frames (predict.py:113-134,
    executed verbatim by tests/golden/make_frames_golden.py), incl. values on / next to every 8-bit bin edge."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "frames_u8_ref.npz"))
    for C in (1, 3):
        u8 = O.frames_to_uint8(z['video_c%d' % C])
        png = z['png_c%d' % C]
        assert np.array_equal(u8[..., 0] if C == 1 else u8, png)
