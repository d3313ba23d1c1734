"""GPU parity of the HBM-bound kernels (ConvLSTM gates, bilinear warp, SloMo fusions) against the
float64 oracle; integer parts (floor / in-bounds corner selection) bit-exact."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import TOL, assert_close, to_cuda

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,F,h,w", [(1, 256, 16, 16), (2, 8, 5, 7), (3, 1, 1, 1), (2, 256, 30, 40)])
def test_gates_forward_backward(cuda, B, F, h, w):
    from video_frame_inpainting_b200 import ops
    rng = np.random.default_rng(20)
    conv = rng.normal(0, 1, (B, 4 * F, h, w)).astype(np.float32)
    state = rng.normal(0, 1, (B, 2 * F, h, w)).astype(np.float32)
    g = rng.normal(0, 1, (B, 2 * F, h, w)).astype(np.float32)
    tc, ts, tg = to_cuda(conv, state, g)
    ns = ops.convlstm_gates_forward(tc, ts, 1.0).cpu().numpy()
    _, ref_ns = O.convlstm_gates(conv, state, 1.0)
    assert_close(ns, ref_ns, what="gates fwd")
    gc, gs = ops.convlstm_gates_backward(tc, ts, tg, 1.0)
    ref_gc, ref_gs = O.convlstm_gates_backward(conv, state, g, 1.0)
    assert_close(gc.cpu().numpy(), ref_gc, what="gates g_conv")
    assert_close(gs.cpu().numpy(), ref_gs, what="gates g_state")
    assert np.all(gs.cpu().numpy()[:, F:] == 0)  # h only feeds the convolution (mcnet.py:288)


def _flow(rng, B, H, W):
    uv = rng.normal(0, 2, (B, 2, H, W)).astype(np.float32)
    mask = rng.uniform(size=(B, 1, H, W)) < 0.08  # >= 5 % of samples pushed far out of bounds
    return np.where(mask, uv * 40, uv).astype(np.float32)


@pytest.mark.parametrize("B,C,H,W", [(1, 3, 32, 40), (2, 1, 7, 5), (1, 3, 256, 320)])
def test_flow_warp_forward(cuda, B, C, H, W):
    import torch
    from video_frame_inpainting_b200 import ops
    rng = np.random.default_rng(21)
    img = rng.uniform(-1, 1, (B, C, H, W)).astype(np.float32)
    uv = _flow(rng, B, H, W)
    ti, tu = to_cuda(img, uv)
    out = ops.flow_warp_forward(ti, tu).cpu().numpy()
    assert_close(out, O.flow_warp(img, uv, coords="f32"), tol=1e-5, what="warp (f32-coordinate twin)")
    # float64 coordinates can fall on the other side of an integer; bilinear is continuous there
    assert_close(out, O.flow_warp(img, uv), tol=2e-3, what="warp (float64 coordinates)")
    # Integer table, bit-exact: the bilinear value is continuous across an integer boundary, so a floor() that is off
    # by one is invisible in `out`.  The flow gradient is not: with img = x^2 the x-derivative of the interpolant is
    # (x0+1)^2 - x0^2 = 2*x0 + 1 wherever the four taps are inside, so floor(ix) is recovered exactly from
    # flow_warp_backward and compared with the oracle's FP32 table (same for y).
    ix, iy, x0, y0 = O.flow_warp_coords_f32(uv)
    inside = (x0 >= 0) & (x0 + 1 < W) & (y0 >= 0) & (y0 + 1 < H)
    assert inside.mean() > 0.5
    ones = torch.ones(B, 1, H, W, device=ti.device)
    sq_x = torch.arange(W, device=ti.device, dtype=torch.float32).pow(2).view(1, 1, 1, W).expand(B, 1, H, W).contiguous()
    sq_y = torch.arange(H, device=ti.device, dtype=torch.float32).pow(2).view(1, 1, H, 1).expand(B, 1, H, W).contiguous()
    gx = ops.flow_warp_backward(sq_x, tu, ones, need_img=False)[1][:, 0].cpu().numpy().astype(np.float64)
    gy = ops.flow_warp_backward(sq_y, tu, ones, need_img=False)[1][:, 1].cpu().numpy().astype(np.float64)
    rec_x0 = np.rint((gx * W / (W - 1) - 1) / 2).astype(np.int32)
    rec_y0 = np.rint((gy * H / (H - 1) - 1) / 2).astype(np.int32)
    assert np.array_equal(rec_x0[inside], x0[inside]) and np.array_equal(rec_y0[inside], y0[inside])


def test_flow_warp_integer_corners_bit_exact(cuda):
    """Flows chosen so that the sampling point is an exact pixel centre: output must equal the
    selected input pixel (or 0 outside), bit for bit -- floor / bounds logic of the sampler."""
    import torch
    from video_frame_inpainting_b200 import ops
    B, C, H, W = 1, 1, 16, 16   # powers of two: every step of the coordinate chain is exact in FP32
    rng = np.random.default_rng(22)
    img = rng.uniform(-1, 1, (B, C, H, W)).astype(np.float32)
    # want ix = (x+u)*(W-1)/W to be the integer k  =>  x + u = k*W/(W-1): not exact in general, so
    # use the oracle's FP32 chain to find which flows land exactly on integers
    ks = rng.integers(-3, W + 3, (B, H, W))
    ls = rng.integers(-3, H + 3, (B, H, W))
    u = (ks * W / (W - 1) - np.arange(W)[None, None, :]).astype(np.float32)
    v = (ls * H / (H - 1) - np.arange(H)[None, :, None]).astype(np.float32)
    uv = np.stack([u, v], 1).astype(np.float32)
    ix, iy, x0, y0 = O.flow_warp_coords_f32(uv)
    exact = (ix == x0) & (iy == y0)
    assert exact.mean() > 0.2
    ti, tu = to_cuda(img, uv)
    out = ops.flow_warp_forward(ti, tu).cpu().numpy()[0, 0]
    inb = (x0 >= 0) & (x0 < W) & (y0 >= 0) & (y0 < H)
    expect = np.where(inb[0], img[0, 0][np.clip(y0[0], 0, H - 1), np.clip(x0[0], 0, W - 1)], 0.0)
    sel = exact[0]
    assert np.array_equal(out[sel], expect[sel].astype(np.float32))


@pytest.mark.parametrize("B,C,H,W", [(1, 3, 32, 40), (2, 1, 7, 5)])
def test_flow_warp_backward(cuda, B, C, H, W):
    from video_frame_inpainting_b200 import ops
    rng = np.random.default_rng(23)
    img = rng.uniform(-1, 1, (B, C, H, W)).astype(np.float32)
    uv = _flow(rng, B, H, W)
    # keep sampling points away from integer coordinates (the gradient is discontinuous there)
    ix, iy, _, _ = O.flow_warp_coords_f32(uv)
    near = (np.abs(ix - np.round(ix)) < 1e-3) | (np.abs(iy - np.round(iy)) < 1e-3)
    uv[:, 0][near] += 0.37
    uv[:, 1][near] += 0.41
    g = rng.uniform(-1, 1, (B, C, H, W)).astype(np.float32)
    ti, tu, tg = to_cuda(img, uv, g)
    gi, gu = ops.flow_warp_backward(ti, tu, tg)
    ref_gi, ref_gu = O.flow_warp_backward(img, uv, g)
    assert_close(gi.cpu().numpy(), ref_gi, tol=5 * TOL, what="warp g_img")
    assert_close(gu.cpu().numpy(), ref_gu, tol=5 * TOL, what="warp g_uv")


@pytest.mark.parametrize("B,C,H,W,T", [(1, 3, 32, 64, 3), (2, 1, 16, 24, 5)])
def test_slomo_fused_stages(cuda, B, C, H, W, T):
    from video_frame_inpainting_b200 import ops
    rng = np.random.default_rng(24)
    i0 = rng.uniform(-1, 1, (B, C, H, W)).astype(np.float32)
    i1 = rng.uniform(-1, 1, (B, C, H, W)).astype(np.float32)
    f01 = np.tanh(rng.normal(0, 1, (B, 2, H, W))).astype(np.float32)
    f10 = np.tanh(rng.normal(0, 1, (B, 2, H, W))).astype(np.float32)
    d0 = np.tanh(rng.normal(0, 1, (B, 2, H, W))).astype(np.float32)
    d1 = np.tanh(rng.normal(0, 1, (B, 2, H, W))).astype(np.float32)
    v0 = rng.uniform(0.05, 0.95, (B, 1, H, W)).astype(np.float32)
    t_i0, t_i1, t_f01, t_f10, t_d0, t_d1, t_v0 = to_cuda(i0, i1, f01, f10, d0, d1, v0)
    for t_ in range(T):
        t = (t_ + 1) / (T + 1)
        ft0, ft1, g0, g1 = ops.slomo_flow_combine_warp(t_i0, t_i1, t_f01, t_f10, t)
        r_ft0, r_ft1 = O.slomo_flow_combine(f01, f10, t)
        assert_close(ft0.cpu().numpy(), r_ft0, what="F_t_0")
        assert_close(ft1.cpu().numpy(), r_ft1, what="F_t_1")
        assert_close(g0.cpu().numpy(), O.flow_warp(i0, ft0.cpu().numpy()), tol=2e-3, what="g0")
        assert_close(g1.cpu().numpy(), O.flow_warp(i1, ft1.cpu().numpy()), tol=2e-3, what="g1")
        out = ops.slomo_refine_blend(t_i0, t_i1, ft0, ft1, t_d0, t_d1, t_v0, t).cpu().numpy()
        ref = O.slomo_refine_blend(i0, i1, ft0.cpu().numpy(), ft1.cpu().numpy(), d0, d1, v0, t)
        assert_close(out, ref, tol=2e-3, what="refine+blend")


@pytest.mark.parametrize("B,C,H,W,T", [(1, 3, 32, 64, 3), (2, 1, 16, 24, 5), (2, 2, 12, 20, 1),
                                       (2, 3, 20, 40, 2),     # four-pixel kernels: ragged tiles in x and in y
                                       (1, 3, 10, 18, 2)])    # W % 4 != 0: per-pixel kernels
def test_slomo_time_batched_stages_forward_and_backward(cuda, B, C, H, W, T):
    """slomo.py:307-340 as two launches over all T middle frames (slomo_interp_input_*, slomo_refine_blend_batched_*):
    values against the float64 oracle, BIT-EXACT against the per-t kernels (same arithmetic, other launch shape),
    gradients against the oracle's adjoints."""
    import torch
    from video_frame_inpainting_b200 import ops
    rng = np.random.default_rng(31)
    # smooth frames: the flow gradient of a warp is a finite difference of the image
    def smooth(n, c):
        low = torch.from_numpy(rng.uniform(-1, 1, (n, c, 4, 6)).astype(np.float32))
        return torch.nn.functional.interpolate(low, size=(H, W), mode='bilinear', align_corners=True).numpy()
    i0, i1 = smooth(B, C), smooth(B, C)
    f01 = np.tanh(rng.normal(0, 1, (B, 2, H, W))).astype(np.float32)
    f10 = np.tanh(rng.normal(0, 1, (B, 2, H, W))).astype(np.float32)
    d0 = np.tanh(rng.normal(0, 1, (T * B, 2, H, W))).astype(np.float32)     # some sums leave [-1, 1]: clamp mask
    d1 = np.tanh(rng.normal(0, 1, (T * B, 2, H, W))).astype(np.float32)
    v0 = rng.uniform(0.05, 0.95, (T * B, 1, H, W)).astype(np.float32)
    t_i0, t_i1, t_f01, t_f10, t_d0, t_d1, t_v0 = to_cuda(i0, i1, f01, f10, d0, d1, v0)
    for t in (t_f01, t_f10, t_d0, t_d1, t_v0):
        t.requires_grad_(True)
    X, c0, c1 = ops.SlomoInterpInputFunction.apply(t_i0, t_i1, t_f01, t_f10, T)
    pred = ops.SlomoRefineBlendFunction.apply(t_i0, t_i1, c0, c1, t_d0, t_d1, t_v0, T)
    rX, rc0, rc1 = O.slomo_interp_input(i0, i1, f01, f10, T)
    assert_close(c0.detach().cpu().numpy(), rc0, what="F_t_0 collector")
    assert_close(c1.detach().cpu().numpy(), rc1, what="F_t_1 collector")
    assert_close(X.detach().cpu().numpy(), rX, tol=2e-3, what="interp_input")
    c0n, c1n = c0.detach().cpu().numpy(), c1.detach().cpu().numpy()
    assert_close(pred.detach().cpu().numpy(), O.slomo_refine_blend_batched(i0, i1, c0n, c1n, d0, d1, v0, T), tol=2e-3,
                 what="pred")
    # the per-t kernels compute the same values bit for bit
    with torch.no_grad():
        for t_ in range(T):
            t = (t_ + 1) / (T + 1)
            ft0, ft1, g0, g1 = ops.slomo_flow_combine_warp(t_i0, t_i1, t_f01, t_f10, t)
            Xt = X[t_ * B:(t_ + 1) * B]
            assert torch.equal(Xt, torch.cat((t_i0, g0, ft0, ft1, g1, t_i1), 1))
            assert torch.equal(c0[:, T - 1 - t_], ft0) and torch.equal(c1[:, T - 1 - t_], ft1)
            out = ops.slomo_refine_blend(t_i0, t_i1, ft0, ft1, t_d0[t_ * B:(t_ + 1) * B].contiguous(),
                                         t_d1[t_ * B:(t_ + 1) * B].contiguous(), t_v0[t_ * B:(t_ + 1) * B].contiguous(), t)
            assert torch.equal(pred[:, T - 1 - t_], out)
    # adjoints
    gX = rng.normal(0, 1, X.shape).astype(np.float32)
    gc0 = rng.normal(0, 1, c0.shape).astype(np.float32)
    gc1 = rng.normal(0, 1, c1.shape).astype(np.float32)
    gp = rng.normal(0, 1, pred.shape).astype(np.float32)
    t_gX, t_gc0, t_gc1, t_gp = to_cuda(gX, gc0, gc1, gp)
    ((X * t_gX).sum() + (c0 * t_gc0).sum() + (c1 * t_gc1).sum() + (pred * t_gp).sum()).backward()
    r_gc0, r_gc1, r_gd0, r_gd1, r_gv0 = O.slomo_refine_blend_batched_backward(i0, i1, c0n, c1n, d0, d1, v0, T, gp)
    assert_close(t_d0.grad.cpu().numpy(), r_gd0, tol=2e-3, what="g dF_t_0")
    assert_close(t_d1.grad.cpu().numpy(), r_gd1, tol=2e-3, what="g dF_t_1")
    assert_close(t_v0.grad.cpu().numpy(), r_gv0, tol=2e-3, what="g V_t_0")
    r_g01, r_g10 = O.slomo_interp_input_backward(i0, i1, f01, f10, T, gX, gc0 + r_gc0, gc1 + r_gc1)
    assert_close(t_f01.grad.cpu().numpy(), r_g01, tol=2e-3, what="g F_0_1")
    assert_close(t_f10.grad.cpu().numpy(), r_g10, tol=2e-3, what="g F_1_0")


def test_slomo_batched_time_equals_per_t_loop(cuda):
    """SloMo.forward with batch_time (one pass over T*B samples) against the reference's per-t loop structure on
    the same weights: same kernels' arithmetic per sample, the convolutions see another batch size."""
    import torch
    from video_frame_inpainting_b200.models.slomo.slomo import SloMoFillInModel
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(5)
    model = SloMoFillInModel(gf_dim=4, c_input_dim=3).cuda()
    pre, fol = torch.rand(2, 2, 3, 32, 64, device='cuda') * 2 - 1, torch.rand(2, 2, 3, 32, 64, device='cuda') * 2 - 1
    outs = {}
    for flag in (True, False):
        model.generator.batch_time = flag
        model.zero_grad()
        out = model(3, pre, fol)
        out['pred'].pow(2).mean().backward()
        outs[flag] = ({k: v.detach().clone() for k, v in out.items()},
                      model.generator.compute_enc.enc1[0].weight.grad.clone())
    for k in outs[True][0]:
        assert O.rel_err(outs[True][0][k].cpu().numpy(), outs[False][0][k].cpu().numpy()) < 1e-4, k
    assert O.rel_err(outs[True][1].cpu().numpy(), outs[False][1].cpu().numpy()) < 2e-3


def test_slomo_stage_kernels_at_config_d_shape(cuda):
    """BASELINE config D's launch shape ([8,3,256,320], T = 3), where the float64 oracle is too slow: the time-batched
    kernels (four pixels per thread, frames staged in shared memory) must reproduce the per-t, per-pixel kernels BIT
    FOR BIT, forward and adjoint, for flows that exercise every clamp branch and the zero padding at all four borders."""
    import torch
    from video_frame_inpainting_b200 import ops
    B, C, H, W, T = 8, 3, 256, 320, 3
    g = torch.Generator(device='cuda').manual_seed(3)
    R = lambda *s: torch.rand(*s, device='cuda', generator=g)
    N = lambda *s: torch.randn(*s, device='cuda', generator=g)
    i0, i1 = R(B, C, H, W), R(B, C, H, W)
    f01, f10 = (N(B, 2, H, W) * 3).requires_grad_(True), (N(B, 2, H, W) * 3).requires_grad_(True)
    d0, d1 = (N(T * B, 2, H, W) * 2).requires_grad_(True), (N(T * B, 2, H, W) * 2).requires_grad_(True)
    v0 = (R(T * B, 1, H, W) * 0.9 + 0.05).requires_grad_(True)
    X, c0, c1 = ops.SlomoInterpInputFunction.apply(i0, i1, f01, f10, T)
    pred = ops.SlomoRefineBlendFunction.apply(i0, i1, c0, c1, d0, d1, v0, T)
    gp = N(*pred.shape)
    (pred * gp).sum().backward()
    with torch.no_grad():
        for t_ in range(T):
            t = (t_ + 1) / (T + 1)
            sl = slice(t_ * B, (t_ + 1) * B)
            ft0, ft1, g0, g1 = ops.slomo_flow_combine_warp(i0, i1, f01, f10, t)
            assert torch.equal(X[sl], torch.cat((i0, g0, ft0, ft1, g1, i1), 1))
            assert torch.equal(c0[:, T - 1 - t_], ft0) and torch.equal(c1[:, T - 1 - t_], ft1)
            out = ops.slomo_refine_blend(i0, i1, ft0, ft1, d0[sl].contiguous(), d1[sl].contiguous(), v0[sl].contiguous(), t)
            assert torch.equal(pred[:, T - 1 - t_], out)
    # adjoint of the refine / blend stage against the composed route (autograd through the per-t torch formulation)
    ft0c, ft1c = c0.detach(), c1.detach()
    d0r, d1r, v0r = (x.detach().clone().requires_grad_(True) for x in (d0, d1, v0))
    total = 0
    for t_ in range(T):
        t = (t_ + 1) / (T + 1)
        sl = slice(t_ * B, (t_ + 1) * B)
        r0 = torch.clamp(d0r[sl] + ft0c[:, T - 1 - t_], -1, 1)
        r1 = torch.clamp(d1r[sl] + ft1c[:, T - 1 - t_], -1, 1)
        a0, a1 = ops.FlowWarpFunction.apply(i0, r0), ops.FlowWarpFunction.apply(i1, r1)
        k0, k1 = (1 - t) * v0r[sl], t * (1 - v0r[sl])
        total = total + ((k0 * a0 + k1 * a1) / (k0 + k1) * gp[:, T - 1 - t_]).sum()
    total.backward()
    assert O.rel_err(d0.grad.cpu().numpy(), d0r.grad.cpu().numpy()) < 1e-4
    assert O.rel_err(d1.grad.cpu().numpy(), d1r.grad.cpu().numpy()) < 1e-4
    assert O.rel_err(v0.grad.cpu().numpy(), v0r.grad.cpu().numpy()) < 1e-4


@pytest.mark.parametrize("C", [1, 3])
def test_motion_prologue_kernels_bit_exact(cuda, C):
    """gray_difference_frames / gray_difference_pair (tai.py:67-74; mcnet.py:439-447; util.py:22-41): bit-identical
    to the reference's elementwise chain evaluated op by op in FP32, close to the float64 oracle, adjoint exact."""
    import torch
    from video_frame_inpainting_b200 import ops
    from video_frame_inpainting_b200.util.util import bgr2gray, bgr2gray_batched, inverse_transform
    g = torch.Generator().manual_seed(11)
    B, K, H, W = 3, 5, 20, 36
    frames = (torch.rand(B, K, C, H, W, generator=g) * 2 - 1).cuda()

    def chain(fr):                                    # tai.py:67-68 as the reference writes it
        x = inverse_transform(fr)
        gray = bgr2gray_batched(x) if C == 3 else x
        return gray[:, 1:] - gray[:, :-1]
    for reverse in (False, True):
        ref = chain(torch.flip(frames, dims=[1]) if reverse else frames)
        got = ops.gray_difference_frames(frames, reverse)
        assert got.shape == ref.shape and torch.equal(got, ref)
    x64 = O.inverse_transform(frames.cpu().numpy())
    gray64 = O.bgr2gray(x64, axis=2) if C == 3 else x64
    assert_close(ops.gray_difference_frames(frames).cpu().numpy(), gray64[:, 1:] - gray64[:, :-1], tol=1e-5, what="gray diff")
    a = (torch.rand(B, C, H, W, generator=g) * 2 - 1).cuda().requires_grad_(True)
    b = (torch.rand(B, C, H, W, generator=g) * 2 - 1).cuda().requires_grad_(True)
    gray01 = lambda v: bgr2gray(inverse_transform(v)) if C == 3 else inverse_transform(v)   # mcnet.py:439-447
    ref = gray01(a) - gray01(b)
    got = ops.GrayDiffPairFunction.apply(a, b)
    assert torch.equal(got.detach(), ref.detach())
    w = torch.rand(ref.shape, generator=g).cuda()
    ga_ref, gb_ref = torch.autograd.grad((ref * w).sum(), (a, b))
    ga, gb = torch.autograd.grad((got * w).sum(), (a, b))
    assert O.rel_err(ga.cpu().numpy(), ga_ref.cpu().numpy()) < 1e-6 and O.rel_err(gb.cpu().numpy(), gb_ref.cpu().numpy()) < 1e-6


def test_gather_concat_matches_torch_cat(cuda):
    """gather_concat_* (tai.py:182,195; mcnet.py:79,91,148): bit-identical to torch.cat / slicing / stacking,
    including a channel-slice source, a constant plane, a strided destination, and the adjoint."""
    import torch
    from video_frame_inpainting_b200 import ops
    g = torch.Generator().manual_seed(3)
    B, T, C, H, W = 3, 4, 5, 6, 8
    xs = [torch.randn(2 * B, C, H, W, generator=g).cuda().requires_grad_(True) for _ in range(T)]
    # (forward stream at t | backward stream at T-1-t), the merge of TAIFillInModel._forward_batched
    blocks = []
    for t in range(T):
        blocks += [(t, 0, B, t * B, 1, 0, C, 0.0), (T - 1 - t, B, B, t * B, 1, C, C, 0.0)]
    got = ops.GatherConcatFunction.apply((T * B, 2 * C, tuple(blocks)), *xs)
    ref = torch.cat([torch.cat([xs[t][:B] for t in range(T)], 0), torch.cat([xs[T - 1 - t][B:] for t in range(T)], 0)], 1)
    assert torch.equal(got, ref)
    w = torch.randn(ref.shape, generator=g).cuda()
    g_got = torch.autograd.grad((got * w).sum(), xs)
    g_ref = torch.autograd.grad((ref * w).sum(), xs)
    assert all(torch.equal(a, b) for a, b in zip(g_got, g_ref))
    # plain channel cat with a channel-slice view as one source, then a constant plane per sample block
    state = torch.randn(B, 2 * C, H, W, generator=g).cuda().requires_grad_(True)
    a = torch.randn(B, 3, H, W, generator=g).cuda().requires_grad_(True)
    hview = state[:, C:]
    got = ops.cat_channels((a, hview))
    ref = torch.cat((a, hview), 1)
    assert torch.equal(got, ref)
    gg = torch.autograd.grad((got * got).sum(), (a, state))
    gr = torch.autograd.grad((ref * ref).sum(), (a, state))
    assert torch.equal(gg[0], gr[0]) and torch.equal(gg[1], gr[1])
    blocks = [(0, 0, B, 0, 1, 0, 3, 0.0), (None, 0, 2, 0, 1, 3, 1, 0.25), (None, 0, 1, 2, 1, 3, 1, 0.75)]
    got = ops.GatherConcatFunction.apply((B, 4, tuple(blocks)), a)
    plane = torch.cat([a.new_full((2, 1, H, W), 0.25), a.new_full((1, 1, H, W), 0.75)], 0)
    assert torch.equal(got, torch.cat((a, plane), 1))
    # stack along dim 1 through a strided destination: out[b, t] = xs[t][b]
    blocks = [(t, 0, 2 * B, t, T, 0, C, 0.0) for t in range(T)]
    got = ops.GatherConcatFunction.apply((2 * B * T, C, tuple(blocks)), *xs).view(2 * B, T, C, H, W)
    assert torch.equal(got, torch.stack(xs, 1))


def test_frames_to_uint8_and_png_layout(cuda, tmp_path):
    """Device-side float -> 8-bit conversion, byte-exact against PNGs written by the reference's own
    save_video_frames (tests/golden/frames_u8_ref.npz), and the predict.py file layout."""
    import os
    import torch
    from PIL import Image
    from video_frame_inpainting_b200 import ops, predict
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "frames_u8_ref.npz"))
    for C in (1, 3):
        video = torch.from_numpy(z['video_c%d' % C]).cuda()
        u8 = ops.frames_to_uint8(video).cpu().numpy()
        png = z['png_c%d' % C]
        assert np.array_equal(u8[..., 0] if C == 1 else u8, png)
        assert np.array_equal(u8, O.frames_to_uint8(z['video_c%d' % C]))
    B, K, T, F_, C, H, W = 2, 2, 3, 2, 3, 16, 24
    g = torch.Generator().manual_seed(3)
    pre, mid, fol = (torch.rand(B, n, C, H + 4, W, generator=g) * 2 - 1 for n in (K, T, F_))
    out = {'pred': mid.cuda() * 0.9, 'pred_forward': mid.cuda() * 0.5}
    files = predict.write_clip_predictions(out, pre, fol, ['vidA_0', 'vidB_7'], str(tmp_path), (H, W),
                                           gt_middle_frames=mid, intermediate_preds=True)
    names = sorted(os.listdir(os.path.join(str(tmp_path), 'vidB_7')))
    assert names == sorted(['gt_preceding_%04d.png' % t for t in range(K)] + ['gt_middle_%04d.png' % (K + t) for t in range(T)]
                           + ['gt_following_%04d.png' % (K + T + t) for t in range(F_)]
                           + ['pred_middle_%04d.png' % (K + t) for t in range(T)]
                           + ['pred_middle_forward_%04d.png' % (K + t) for t in range(T)])
    assert len(files) == 2 * len(names)
    img = np.array(Image.open(os.path.join(str(tmp_path), 'vidA_0', 'pred_middle_%04d.png' % K)))
    assert img.shape == (H, W, 3)
    assert np.array_equal(img, O.frames_to_uint8((mid[0, :1, :, :H, :W] * 0.9).numpy())[0])


@pytest.mark.parametrize("N,C,H,W", [(2, 5, 6, 8), (3, 7, 5, 3), (64, 64, 32, 32), (1, 1, 1, 1), (4, 130, 4, 4)])
@pytest.mark.parametrize("act,alpha", [("relu", 0.0), ("leaky", 0.1), ("none", 0.0)])
def test_bias_act_matches_library_ops(cuda, N, C, H, W, act, alpha):
    """Bias + activation epilogue against the library ops it replaces (broadcast add, relu / leaky_relu, their
    backward and the bias-gradient sum): forward bit-identical, input gradient bit-identical, bias gradient to
    1e-5 (another summation order)."""
    import torch
    import torch.nn.functional as F
    from video_frame_inpainting_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(50)
    y = torch.randn(N, C, H, W, device=cuda, generator=g)
    b = torch.randn(C, device=cuda, generator=g)
    go = torch.randn(N, C, H, W, device=cuda, generator=g)
    y_ref = y.clone().requires_grad_()
    b_ref = b.clone().requires_grad_()
    pre = y_ref + b_ref.view(1, C, 1, 1)
    ref = {"relu": F.relu, "leaky": lambda t: F.leaky_relu(t, alpha), "none": lambda t: t}[act](pre)
    ref.backward(go)
    y2 = (y.clone().requires_grad_() * 1.0)          # a non-leaf the Function may overwrite in place
    b2 = b.clone().requires_grad_()
    src = y2
    out = ops.BiasActFunction.apply(y2, b2, act, alpha)
    assert out.data_ptr() == src.data_ptr()           # in place
    assert torch.equal(out, ref)
    gin, gb = ops.bias_act_backward(go, out.detach(), act, alpha)
    assert torch.equal(gin, y_ref.grad)
    assert_close(gb.cpu().numpy(), b_ref.grad.cpu().numpy(), tol=1e-5, what="bias gradient")
    assert torch.equal(gb, ops.bias_act_backward(go, out.detach(), act, alpha)[1]), "deterministic"


def test_fused_sequential_equals_plain_sequential(cuda):
    """FusedSequential (bias-free convolution + bias/activation pass) against nn.Sequential with the same children:
    same state_dict keys, same outputs, same gradients."""
    import torch
    import torch.nn as nn
    from video_frame_inpainting_b200.models.layers import FusedSequential
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    mods = lambda: [nn.Conv2d(3, 8, 3, padding=1), nn.ReLU(), nn.Conv2d(8, 8, 5, padding=2), nn.LeakyReLU(0.1),
                    nn.ConvTranspose2d(8, 4, 3, padding=1), nn.ReLU(), nn.Conv2d(4, 2, 1), nn.Tanh()]
    plain = nn.Sequential(*mods()).cuda()
    fused = FusedSequential(*mods()).cuda()
    assert list(plain.state_dict().keys()) == list(fused.state_dict().keys())
    fused.load_state_dict(plain.state_dict())
    x = torch.randn(2, 3, 16, 20, device=cuda)
    xa, xb = x.clone().requires_grad_(), x.clone().requires_grad_()
    ya, yb = plain(xa), fused(xb)
    assert_close(yb.detach().cpu().numpy(), ya.detach().cpu().numpy(), tol=1e-5, what="fused sequential fwd")
    g = torch.randn_like(ya)
    ya.backward(g)
    yb.backward(g)
    assert_close(xb.grad.cpu().numpy(), xa.grad.cpu().numpy(), tol=1e-4, what="fused sequential grad x")
    for (n, pa), (_, pb) in zip(plain.named_parameters(), fused.named_parameters()):
        assert_close(pb.grad.cpu().numpy(), pa.grad.cpu().numpy(), tol=1e-4, what="fused sequential grad " + n)


def test_l2_normalize_and_sn_discriminator_route(cuda):
    """One-launch l2 normalisation of the spectral-norm power iteration against the five-op library chain, and the
    discriminator (bias + LeakyReLU epilogue, fused normalisation) against the same module on the CPU."""
    import copy
    import torch
    from video_frame_inpainting_b200 import ops
    from video_frame_inpainting_b200.discriminators.SNDiscriminator import SNDiscriminator
    for n in (1, 7, 64, 513, 8192):
        v = torch.randn(1, n, device=cuda)
        ref = v / (((v ** 2).sum()) ** 0.5 + 1e-12)
        assert_close(ops.l2_normalize(v).cpu().numpy(), ref.cpu().numpy(), tol=1e-6, what="l2 normalize")
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    d_cpu = SNDiscriminator((32, 32), 1, 3, 8, 3)
    d_gpu = copy.deepcopy(d_cpu).cuda()
    x = torch.randn(2, 6, 1, 32, 32)
    for _ in range(2):                     # two calls: the in-place weight normalisation and u carry over
        a = d_cpu(x)
        b = d_gpu(x.cuda())
    assert_close(b.detach().cpu().numpy(), a.detach().numpy(), tol=1e-3, what="SN discriminator logits")
    a.sum().backward()
    b.sum().backward()
    for (n_, pa), (_, pb) in zip(d_cpu.named_parameters(), d_gpu.named_parameters()):
        assert_close(pb.grad.cpu().numpy(), pa.grad.numpy(), tol=2e-3, what="SN discriminator grad " + n_)
