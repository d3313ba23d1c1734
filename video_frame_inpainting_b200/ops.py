"""Tensor-level wrappers over the C ABI (include/tai_b200.h) and their autograd Functions.

torch is used here for device memory, streams and autograd bookkeeping only; all arithmetic of the
hot path happens in libtai_b200.so.  CPU tensors raise NotImplementedError exactly like the reference
operator (src/separable_convolution/SeparableConvolution.py:48-49,86-87) -- there is no fallback.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _check(name, *tensors):
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise NotImplementedError("%s: CPU tensors are not supported (no CPU version, as in the reference)" % name)
        if t.dtype != torch.float32:
            raise TypeError("%s: expected float32 tensors, got %s" % (name, t.dtype))
        assert t.is_contiguous(), "%s: tensors must be contiguous" % name
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise ValueError("%s: tensors live on different devices" % name)
    return dev


# ------------------------------------------------------------------------------------------------
# separable convolution
# ------------------------------------------------------------------------------------------------

def sepconv_shapes(input, vertical, horizontal, ks):
    """The shape algebra and asserts of SeparableConvolution.py:16-33."""
    B, C, Hi, Wi = input.shape
    fs = min(vertical.size(1), horizontal.size(1))
    Ho = min(vertical.size(2), horizontal.size(2))
    Wo = min(vertical.size(3), horizontal.size(3))
    assert Hi - ks == Ho - 1
    assert Wi - ks == Wo - 1
    assert fs == ks
    assert vertical.shape == horizontal.shape == (B, ks, Ho, Wo), \
        "vertical/horizontal must both be [B, ks, Ho, Wo]"
    return B, C, Hi, Wi, Ho, Wo


def sepconv_forward(input, vertical, horizontal, ks):
    dev = _check("sepconv_forward", input, vertical, horizontal)
    B, C, Hi, Wi, Ho, Wo = sepconv_shapes(input, vertical, horizontal, ks)
    with torch.cuda.device(dev):
        out = torch.empty((B, C, Ho, Wo), device=dev, dtype=torch.float32)
        _lib.call("SeparableConvolution_cuda_forward_b200", _ptr(input), _ptr(vertical), _ptr(horizontal),
                  _ptr(out), B, C, Hi, Wi, ks, _stream())
    return out


def sepconv_backward(grad_output, input, vertical, horizontal, ks, needs=(True, True, True)):
    dev = _check("sepconv_backward", grad_output, input, vertical, horizontal)
    B, C, Hi, Wi, Ho, Wo = sepconv_shapes(input, vertical, horizontal, ks)
    assert grad_output.shape == (B, C, Ho, Wo)
    with torch.cuda.device(dev):
        gi = torch.empty_like(input) if needs[0] else None
        gv = torch.empty_like(vertical) if needs[1] else None
        gh = torch.empty_like(horizontal) if needs[2] else None
        if any(needs):
            _lib.call("SeparableConvolution_cuda_backward_b200", _ptr(grad_output), _ptr(input), _ptr(vertical),
                      _ptr(horizontal), _ptr(gi), _ptr(gv), _ptr(gh), B, C, Hi, Wi, ks, _stream())
    return gi, gv, gh


def tai_fused_forward(pred_f, pred_b, v1, h1, v2, h2, ks, a=0.5, b=0.5, emit_intermediate=True):
    """pad + 2 x sepconv + blend (tai.py:229-236,105).  Returns (pred, dot1, dot2); dot1/dot2 are None
    when emit_intermediate is False."""
    dev = _check("tai_fused_forward", pred_f, pred_b, v1, h1, v2, h2)
    B, C, H, W = pred_f.shape
    assert pred_b.shape == pred_f.shape
    for k in (v1, h1, v2, h2):
        assert k.shape == (B, ks, H, W), "kernel maps must be [B, ks, H, W]"
    with torch.cuda.device(dev):
        pred = torch.empty_like(pred_f)
        d1 = torch.empty_like(pred_f) if emit_intermediate else None
        d2 = torch.empty_like(pred_f) if emit_intermediate else None
        _lib.call("tai_fused_forward_b200", _ptr(pred_f), _ptr(pred_b), _ptr(v1), _ptr(h1), _ptr(v2), _ptr(h2),
                  _ptr(pred), _ptr(d1), _ptr(d2), B, C, H, W, ks, float(a), float(b), _stream())
    return pred, d1, d2


def tai_fused_backward(g_pred, g_dot1, g_dot2, pred_f, pred_b, v1, h1, v2, h2, ks, a, b,
                       need_pred=(True, True), need_maps=True):
    dev = _check("tai_fused_backward", g_pred, g_dot1, g_dot2, pred_f, pred_b, v1, h1, v2, h2)
    B, C, H, W = pred_f.shape
    with torch.cuda.device(dev):
        ws_bytes = _lib.load().tai_fused_backward_workspace_bytes(B, C, H, W, ks)
        ws = torch.empty(ws_bytes // 4, device=dev, dtype=torch.float32)
        gpf = torch.empty_like(pred_f) if need_pred[0] else None
        gpb = torch.empty_like(pred_b) if need_pred[1] else None
        gv1, gh1, gv2, gh2 = (torch.empty_like(v1) if need_maps else None for _ in range(4))
        _lib.call("tai_fused_backward_b200", _ptr(g_pred), _ptr(g_dot1), _ptr(g_dot2), _ptr(pred_f), _ptr(pred_b),
                  _ptr(v1), _ptr(h1), _ptr(v2), _ptr(h2), _ptr(gpf), _ptr(gpb), _ptr(gv1), _ptr(gh1), _ptr(gv2),
                  _ptr(gh2), _ptr(ws), B, C, H, W, ks, float(a), float(b), _stream())
    return gpf, gpb, gv1, gh1, gv2, gh2


def replication_pad_forward(x, p):
    dev = _check("replication_pad_forward", x)
    H, W = x.shape[-2:]
    N = x.numel() // (H * W)
    with torch.cuda.device(dev):
        out = torch.empty(x.shape[:-2] + (H + 2 * p, W + 2 * p), device=dev, dtype=torch.float32)
        _lib.call("replication_pad_forward_b200", _ptr(x), _ptr(out), N, H, W, p, _stream())
    return out


def replication_pad_backward(g, p):
    dev = _check("replication_pad_backward", g)
    Hp, Wp = g.shape[-2:]
    H, W = Hp - 2 * p, Wp - 2 * p
    N = g.numel() // (Hp * Wp)
    with torch.cuda.device(dev):
        out = torch.empty(g.shape[:-2] + (H, W), device=dev, dtype=torch.float32)
        _lib.call("replication_pad_backward_b200", _ptr(g), _ptr(out), N, H, W, p, _stream())
    return out


class SeparableConvolutionFunction(torch.autograd.Function):
    """autograd.Function with the reference's contract (SeparableConvolution.py:6-92):
    apply(input, vertical, horizontal, ks) -> output; backward -> (gI, gV, gH, None); gradients are
    not differentiable (no double backward)."""

    @staticmethod
    def forward(ctx, input, vertical, horizontal, ks=51):
        ctx.save_for_backward(input, vertical, horizontal)
        ctx.constant = ks
        return sepconv_forward(input, vertical, horizontal, ks)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_output):
        input, vertical, horizontal = ctx.saved_tensors
        gi, gv, gh = sepconv_backward(grad_output.contiguous(), input, vertical, horizontal, ctx.constant,
                                      needs=tuple(ctx.needs_input_grad[:3]))
        return gi, gv, gh, None


class TAIBlendSepConvFunction(torch.autograd.Function):
    """Fused pad + sepconv x2 + blend.  apply(pred_f, pred_b, v1, h1, v2, h2, ks, a, b) ->
    (pred, dot1, dot2)."""

    @staticmethod
    def forward(ctx, pred_f, pred_b, v1, h1, v2, h2, ks, a, b):
        ctx.save_for_backward(pred_f, pred_b, v1, h1, v2, h2)
        ctx.consts = (ks, float(a), float(b))
        ctx.set_materialize_grads(False)
        return tai_fused_forward(pred_f, pred_b, v1, h1, v2, h2, ks, a, b, emit_intermediate=True)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_pred, g_dot1, g_dot2):
        pred_f, pred_b, v1, h1, v2, h2 = ctx.saved_tensors
        ks, a, b = ctx.consts
        if g_pred is None and g_dot1 is None and g_dot2 is None:
            return (None,) * 9
        cont = lambda g: None if g is None else g.contiguous()
        need = ctx.needs_input_grad
        gpf, gpb, gv1, gh1, gv2, gh2 = tai_fused_backward(
            cont(g_pred), cont(g_dot1), cont(g_dot2), pred_f, pred_b, v1, h1, v2, h2, ks, a, b,
            need_pred=(need[0], need[1]), need_maps=any(need[2:6]))
        return gpf, gpb, gv1, gh1, gv2, gh2, None, None, None


def tai_blend_sepconv(pred_f, pred_b, v1, h1, v2, h2, ks, a=0.5, b=0.5, emit_intermediate=True):
    """Public fused entry.  Differentiable; returns (pred, dot1, dot2)."""
    if torch.is_grad_enabled() and any(t.requires_grad for t in (pred_f, pred_b, v1, h1, v2, h2)):
        return TAIBlendSepConvFunction.apply(pred_f, pred_b, v1, h1, v2, h2, ks, a, b)
    return tai_fused_forward(pred_f, pred_b, v1, h1, v2, h2, ks, a, b, emit_intermediate)


# ------------------------------------------------------------------------------------------------
# ConvLSTM gates
# ------------------------------------------------------------------------------------------------

def convlstm_gates_forward(conv_out, state, forget_bias=1.0):
    dev = _check("convlstm_gates_forward", conv_out, state)
    B, F2 = state.shape[:2]
    assert F2 % 2 == 0 and conv_out.shape[1] == 2 * F2 and conv_out.shape[0] == B
    assert conv_out.shape[2:] == state.shape[2:]
    F = F2 // 2
    HW = state.numel() // (B * F2)
    with torch.cuda.device(dev):
        new_state = torch.empty_like(state)
        _lib.call("convlstm_gates_forward_b200", _ptr(conv_out), _ptr(state), _ptr(new_state), B, F, HW,
                  float(forget_bias), _stream())
    return new_state


def convlstm_gates_backward(conv_out, state, g_new_state, forget_bias=1.0):
    dev = _check("convlstm_gates_backward", conv_out, state, g_new_state)
    B, F2 = state.shape[:2]
    F = F2 // 2
    HW = state.numel() // (B * F2)
    with torch.cuda.device(dev):
        g_conv = torch.empty_like(conv_out)
        g_state = torch.empty_like(state)
        _lib.call("convlstm_gates_backward_b200", _ptr(conv_out), _ptr(state), _ptr(g_new_state), _ptr(g_conv),
                  _ptr(g_state), B, F, HW, float(forget_bias), _stream())
    return g_conv, g_state


class ConvLstmGatesFunction(torch.autograd.Function):
    """apply(conv_out, state, forget_bias) -> new_state = cat(c', h')   (mcnet.py:287-293)."""

    @staticmethod
    def forward(ctx, conv_out, state, forget_bias):
        ctx.save_for_backward(conv_out, state)
        ctx.forget_bias = float(forget_bias)
        return convlstm_gates_forward(conv_out, state, forget_bias)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_new_state):
        conv_out, state = ctx.saved_tensors
        g_conv, g_state = convlstm_gates_backward(conv_out, state, g_new_state.contiguous(), ctx.forget_bias)
        return g_conv, g_state, None


# ------------------------------------------------------------------------------------------------
# decoder resampling
# ------------------------------------------------------------------------------------------------

def upsample_bilinear2x_forward(x):
    dev = _check("upsample_bilinear2x_forward", x)
    B, C, H, W = x.shape
    with torch.cuda.device(dev):
        out = torch.empty(B, C, 2 * H, 2 * W, device=dev, dtype=torch.float32)
        _lib.call("upsample_bilinear2x_forward_b200", _ptr(x), _ptr(out), B * C, H, W, _stream())
    return out


def upsample_bilinear2x_backward(grad_out):
    dev = _check("upsample_bilinear2x_backward", grad_out)
    B, C, Ho, Wo = grad_out.shape
    assert Ho % 2 == 0 and Wo % 2 == 0
    with torch.cuda.device(dev):
        gin = torch.empty(B, C, Ho // 2, Wo // 2, device=dev, dtype=torch.float32)
        _lib.call("upsample_bilinear2x_backward_b200", _ptr(grad_out), _ptr(gin), B * C, Ho // 2, Wo // 2, _stream())
    return gin


class UpsampleBilinear2xFunction(torch.autograd.Function):
    """apply(x[B,C,H,W]) -> [B,C,2H,2W]: nn.Upsample(scale_factor=2, mode='bilinear') with the torch-0.3.1
    (align-corners) mapping   (tai.py:283,337,343; slomo.py:113-149)."""

    @staticmethod
    def forward(ctx, x):
        return upsample_bilinear2x_forward(x)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        return upsample_bilinear2x_backward(grad_out.contiguous())


def unpool_add_forward(x, res):
    dev = _check("unpool_add_forward", x, res)
    B, C, H, W = x.shape
    assert res.shape == (B, C, 2 * H, 2 * W), "residual must be [B, C, 2H, 2W]"
    with torch.cuda.device(dev):
        out = torch.empty_like(res)
        _lib.call("unpool_add_forward_b200", _ptr(x), _ptr(res), _ptr(out), B * C, H, W, _stream())
    return out


def unpool_backward(grad_out):
    dev = _check("unpool_backward", grad_out)
    B, C, Ho, Wo = grad_out.shape
    with torch.cuda.device(dev):
        gx = torch.empty(B, C, Ho // 2, Wo // 2, device=dev, dtype=torch.float32)
        _lib.call("unpool_backward_b200", _ptr(grad_out), _ptr(gx), B * C, Ho // 2, Wo // 2, _stream())
    return gx


class UnpoolAddFunction(torch.autograd.Function):
    """apply(x[B,C,H,W], res[B,C,2H,2W]) -> fixed_unpooling(x) + res   (mcnet.py:234-236,240-256)."""

    @staticmethod
    def forward(ctx, x, res):
        ctx.needs = (ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return unpool_add_forward(x, res)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        grad_out = grad_out.contiguous()
        gx = unpool_backward(grad_out) if ctx.needs[0] else None
        return gx, (grad_out if ctx.needs[1] else None)


def maxpool2x2_forward(x):
    """nn.MaxPool2d(2) (mcnet.py:28-45; slomo.py:47-85): returns (out, code) -- code is the 2-bit position of
    the selected element, one byte per output element (the library keeps an int64 flat index)."""
    dev = _check("maxpool2x2_forward", x)
    H, W = x.shape[-2:]
    assert H >= 2 and W >= 2
    planes = x.numel() // (H * W)
    with torch.cuda.device(dev):
        out = torch.empty(x.shape[:-2] + (H // 2, W // 2), device=dev, dtype=torch.float32)
        code = torch.empty(out.shape, device=dev, dtype=torch.uint8)
        _lib.call("maxpool2x2_forward_b200", _ptr(x), _ptr(out), _ptr(code), planes, H, W, _stream())
    return out, code


def maxpool2x2_backward(grad_out, code, H, W):
    dev = _check("maxpool2x2_backward", grad_out)
    assert code.is_cuda and code.dtype == torch.uint8 and code.is_contiguous() and code.shape == grad_out.shape
    planes = grad_out.numel() // (grad_out.shape[-2] * grad_out.shape[-1])
    with torch.cuda.device(dev):
        gin = torch.empty(grad_out.shape[:-2] + (H, W), device=dev, dtype=torch.float32)
        _lib.call("maxpool2x2_backward_b200", _ptr(grad_out), _ptr(code), _ptr(gin), planes, H, W, _stream())
    return gin


class MaxPool2x2Function(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        out, code = maxpool2x2_forward(x)
        ctx.save_for_backward(code)
        ctx.hw = x.shape[-2:]
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        (code,) = ctx.saved_tensors
        return maxpool2x2_backward(grad_out.contiguous(), code, *ctx.hw)


# ------------------------------------------------------------------------------------------------
# Super SloMo warp / blend
# ------------------------------------------------------------------------------------------------

def flow_warp_forward(img, uv):
    dev = _check("flow_warp_forward", img, uv)
    B, C, H, W = img.shape
    assert uv.shape == (B, 2, H, W)
    with torch.cuda.device(dev):
        out = torch.empty_like(img)
        _lib.call("flow_warp_forward_b200", _ptr(img), _ptr(uv), _ptr(out), B, C, H, W, _stream())
    return out


def flow_warp_backward(img, uv, grad_out, need_img=True, need_uv=True):
    dev = _check("flow_warp_backward", img, uv, grad_out)
    B, C, H, W = img.shape
    with torch.cuda.device(dev):
        g_img = torch.empty_like(img) if need_img else None
        g_uv = torch.empty_like(uv) if need_uv else None
        if need_img or need_uv:
            _lib.call("flow_warp_backward_b200", _ptr(img), _ptr(uv), _ptr(grad_out), _ptr(g_img), _ptr(g_uv),
                      B, C, H, W, _stream())
    return g_img, g_uv


class FlowWarpFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, uv):
        ctx.save_for_backward(img, uv)
        return flow_warp_forward(img, uv)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        img, uv = ctx.saved_tensors
        return flow_warp_backward(img, uv, grad_out.contiguous(), ctx.needs_input_grad[0], ctx.needs_input_grad[1])


def slomo_flow_combine_warp(i0, i1, f01, f10, t):
    """slomo.py:312-316 fused: returns (F_t_0, F_t_1, g_I0_F_t_0, g_I1_F_t_1).  Forward only."""
    dev = _check("slomo_flow_combine_warp", i0, i1, f01, f10)
    B, C, H, W = i0.shape
    assert i1.shape == i0.shape and f01.shape == f10.shape == (B, 2, H, W)
    with torch.cuda.device(dev):
        ft0, ft1 = torch.empty_like(f01), torch.empty_like(f01)
        g0, g1 = torch.empty_like(i0), torch.empty_like(i0)
        _lib.call("slomo_flow_combine_warp_forward_b200", _ptr(i0), _ptr(i1), _ptr(f01), _ptr(f10), float(t),
                  _ptr(ft0), _ptr(ft1), _ptr(g0), _ptr(g1), B, C, H, W, _stream())
    return ft0, ft1, g0, g1


def slomo_refine_blend(i0, i1, f_t0, f_t1, d_t0, d_t1, v_t0, t):
    """slomo.py:320-328 fused: returns the interpolated frame.  Forward only."""
    dev = _check("slomo_refine_blend", i0, i1, f_t0, f_t1, d_t0, d_t1, v_t0)
    B, C, H, W = i0.shape
    assert v_t0.shape == (B, 1, H, W)
    with torch.cuda.device(dev):
        out = torch.empty_like(i0)
        _lib.call("slomo_refine_blend_forward_b200", _ptr(i0), _ptr(i1), _ptr(f_t0), _ptr(f_t1), _ptr(d_t0),
                  _ptr(d_t1), _ptr(v_t0), float(t), _ptr(out), B, C, H, W, _stream())
    return out


class SlomoInterpInputFunction(torch.autograd.Function):
    """slomo.py:312-318 for all T middle frames at once: ``apply(I0, I1, F_0_1, F_1_0, T)`` ->
    (interp_input [T*B,4C+4,H,W] in sample order n = t*B + b, F_t_0_collector, F_t_1_collector [B,T,2,H,W] in the
    reference's reversed time order).  Differentiable w.r.t. the two flows (gather-only adjoint); I0 / I1 must not
    require gradients (they are network inputs; SloMo.forward takes the composed route otherwise)."""

    @staticmethod
    def forward(ctx, i0, i1, f01, f10, T):
        dev = _check("slomo_interp_input", i0, i1, f01, f10)
        B, C, H, W = i0.shape
        assert i1.shape == i0.shape and f01.shape == f10.shape == (B, 2, H, W)
        with torch.cuda.device(dev):
            x = torch.empty(T * B, 4 * C + 4, H, W, device=i0.device, dtype=i0.dtype)
            ft0c = torch.empty(B, T, 2, H, W, device=i0.device, dtype=i0.dtype)
            ft1c = torch.empty_like(ft0c)
            _lib.call("slomo_interp_input_forward_b200", _ptr(i0), _ptr(i1), _ptr(f01), _ptr(f10), _ptr(x), _ptr(ft0c),
                      _ptr(ft1c), B, T, C, H, W, _stream())
        ctx.save_for_backward(i0, i1, f01, f10)
        ctx.T = T
        return x, ft0c, ft1c

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gx, gft0c, gft1c):
        i0, i1, f01, f10 = ctx.saved_tensors
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            raise NotImplementedError("SlomoInterpInputFunction: no gradient w.r.t. the frames I0 / I1")
        B, C, H, W = i0.shape
        with torch.cuda.device(i0.device):
            gx = gx.contiguous()
            gft0c = gft0c.contiguous() if gft0c is not None else None
            gft1c = gft1c.contiguous() if gft1c is not None else None
            g01, g10 = torch.empty_like(f01), torch.empty_like(f10)
            _lib.call("slomo_interp_input_backward_b200", _ptr(i0), _ptr(i1), _ptr(f01), _ptr(f10), _ptr(gx),
                      _ptr(gft0c), _ptr(gft1c), _ptr(g01), _ptr(g10), B, ctx.T, C, H, W, _stream())
        return None, None, g01, g10, None


class SlomoRefineBlendFunction(torch.autograd.Function):
    """slomo.py:320-328 for all T middle frames at once: ``apply(I0, I1, F_t_0_collector, F_t_1_collector, dF_t_0,
    dF_t_1, V_t_0, T)`` -> pred [B,T,C,H,W] (reversed time order).  dF_t_*, V_t_0 in sample order n = t*B + b.
    Differentiable w.r.t. everything but the frames."""

    @staticmethod
    def forward(ctx, i0, i1, ft0c, ft1c, d0, d1, v0, T):
        dev = _check("slomo_refine_blend_batched", i0, i1, ft0c, ft1c, d0, d1, v0)
        B, C, H, W = i0.shape
        assert ft0c.shape == ft1c.shape == (B, T, 2, H, W) and d0.shape == d1.shape == (T * B, 2, H, W)
        assert v0.shape == (T * B, 1, H, W)
        with torch.cuda.device(dev):
            pred = torch.empty(B, T, C, H, W, device=i0.device, dtype=i0.dtype)
            _lib.call("slomo_refine_blend_batched_forward_b200", _ptr(i0), _ptr(i1), _ptr(ft0c), _ptr(ft1c), _ptr(d0),
                      _ptr(d1), _ptr(v0), _ptr(pred), B, T, C, H, W, _stream())
        ctx.save_for_backward(i0, i1, ft0c, ft1c, d0, d1, v0)
        ctx.T = T
        return pred

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gpred):
        i0, i1, ft0c, ft1c, d0, d1, v0 = ctx.saved_tensors
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            raise NotImplementedError("SlomoRefineBlendFunction: no gradient w.r.t. the frames I0 / I1")
        B, C, H, W = i0.shape
        with torch.cuda.device(i0.device):
            gpred = gpred.contiguous()
            gft0c, gft1c = torch.empty_like(ft0c), torch.empty_like(ft1c)
            gd0, gd1, gv0 = torch.empty_like(d0), torch.empty_like(d1), torch.empty_like(v0)
            _lib.call("slomo_refine_blend_batched_backward_b200", _ptr(i0), _ptr(i1), _ptr(ft0c), _ptr(ft1c), _ptr(d0),
                      _ptr(d1), _ptr(v0), _ptr(gpred), _ptr(gft0c), _ptr(gft1c), _ptr(gd0), _ptr(gd1), _ptr(gv0),
                      B, ctx.T, C, H, W, _stream())
        return None, None, gft0c, gft1c, gd0, gd1, gv0, None


# ------------------------------------------------------------------------------------------------
# reconstruction losses (MSELoss + GDL in one pass)
# ------------------------------------------------------------------------------------------------

def l2_gdl_loss_forward(pred, target, add=1.0, mul=0.5):
    """(mse, gdl) of v01 = (v + add) * mul as a 2-element tensor   (environments.py:363-371; losses.py:24-45)."""
    dev = _check("l2_gdl_loss_forward", pred, target)
    assert pred.shape == target.shape and pred.dim() >= 2
    H, W = pred.shape[-2:]
    planes = pred.numel() // (H * W)
    with torch.cuda.device(dev):
        out = torch.empty(2, device=dev, dtype=torch.float32)
        nbytes = int(_lib.load().l2_gdl_loss_workspace_bytes(planes, H, W))
        ws = torch.empty(max(nbytes, 8) // 4, device=dev, dtype=torch.float32)
        _lib.call("l2_gdl_loss_forward_b200", _ptr(pred), _ptr(target), planes, H, W, float(add), float(mul),
                  _ptr(out), _ptr(ws), _stream())
    return out


def l2_gdl_loss_backward(pred, target, grad_mse, grad_gdl, add=1.0, mul=0.5):
    """grad_mse / grad_gdl: 1-element CUDA tensors (or None) -- read on the device."""
    dev = _check("l2_gdl_loss_backward", pred, target, grad_mse, grad_gdl)
    H, W = pred.shape[-2:]
    planes = pred.numel() // (H * W)
    with torch.cuda.device(dev):
        g = torch.empty_like(pred)
        _lib.call("l2_gdl_loss_backward_b200", _ptr(pred), _ptr(target), planes, H, W, float(add), float(mul),
                  _ptr(grad_mse), _ptr(grad_gdl), _ptr(g), _stream())
    return g


class L2GDLLossFunction(torch.autograd.Function):
    """(pred, target) -> (mse, gdl), both 0-dim.  The target is treated as a constant, as in the training
    step (ground-truth frames).  Backward recomputes the terms from pred / target: nothing but the two
    input tensors is kept."""

    @staticmethod
    def forward(ctx, pred, target, add, mul):
        pred = pred.contiguous()
        target = target.contiguous()
        out = l2_gdl_loss_forward(pred, target, add, mul)
        ctx.save_for_backward(pred, target)
        ctx.affine = (add, mul)
        return out[0], out[1]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_mse, g_gdl):
        pred, target = ctx.saved_tensors
        if not ctx.needs_input_grad[0]:
            return None, None, None, None
        g_mse = g_mse.reshape(1).contiguous() if g_mse is not None else None
        g_gdl = g_gdl.reshape(1).contiguous() if g_gdl is not None else None
        return l2_gdl_loss_backward(pred, target, g_mse, g_gdl, *ctx.affine), None, None, None


def l2_gdl_loss(pred, target, add=1.0, mul=0.5):
    return L2GDLLossFunction.apply(pred, target, add, mul)


# ------------------------------------------------------------------------------------------------
# bias + activation epilogue of the convolution layers
# ------------------------------------------------------------------------------------------------

ACT_CODES = {"none": 0, "relu": 1, "leaky": 2}


def bias_act_forward_(y, bias, act="relu", alpha=0.0):
    """In place: y[N,C,...] = act(y + bias[c])."""
    dev = _check("bias_act_forward_", y, bias)
    N, C = y.shape[0], y.shape[1]
    assert bias.shape == (C,)
    with torch.cuda.device(dev):
        _lib.call("bias_act_forward_b200", _ptr(y), _ptr(bias), N, C, y.numel() // (N * C), ACT_CODES[act], float(alpha),
                  _stream())
    return y


def bias_act_backward(grad_out, out, act="relu", alpha=0.0):
    """-> (grad_in, grad_bias); grad_in is grad_out itself for act == 'none'."""
    dev = _check("bias_act_backward", grad_out, out)
    N, C = grad_out.shape[0], grad_out.shape[1]
    code = ACT_CODES[act]
    with torch.cuda.device(dev):
        gb = torch.empty(C, device=dev, dtype=torch.float32)
        ws = torch.empty(max(int(_lib.load().bias_act_backward_workspace_bytes(N, C)), 4) // 4, device=dev,
                         dtype=torch.float32)
        gin = torch.empty_like(grad_out) if code != 0 else None
        _lib.call("bias_act_backward_b200", _ptr(grad_out), _ptr(out), _ptr(gin), _ptr(gb), _ptr(ws), N, C,
                  grad_out.numel() // (N * C), code, float(alpha), _stream())
    return (gin if gin is not None else grad_out), gb


class BiasActFunction(torch.autograd.Function):
    """(y, bias) -> act(y + bias) IN PLACE on y, which must be the fresh output of a bias-free convolution (nothing
    else may hold it: the convolution's own backward needs its input and weight only)."""

    @staticmethod
    def forward(ctx, y, bias, act, alpha):
        assert y.is_contiguous()
        bias_act_forward_(y, bias.contiguous(), act, alpha)
        ctx.mark_dirty(y)
        ctx.act, ctx.alpha = act, alpha
        if act != "none":
            ctx.save_for_backward(y)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        out = ctx.saved_tensors[0] if ctx.act != "none" else None
        gin, gb = bias_act_backward(grad_out.contiguous(), out, ctx.act, ctx.alpha)
        return gin, gb, None, None


def l2_normalize(v, eps=1e-12):
    """v / (||v||_2 + eps) in one launch (SNDiscriminator.py:5-7); no autograd (the power iteration runs without)."""
    dev = _check("l2_normalize", v)
    with torch.cuda.device(dev):
        out = torch.empty_like(v)
        _lib.call("l2_normalize_b200", _ptr(v), _ptr(out), v.numel(), float(eps), _stream())
    return out


# ------------------------------------------------------------------------------------------------
# gather-concatenation (tai.py:182,195; mcnet.py:79,91,148; the stack-then-cat of the batched kernel network)
# ------------------------------------------------------------------------------------------------

class GatherConcatFunction(torch.autograd.Function):
    """``apply(spec, *tensors)`` -> dst [N, Ctot, H, W].  ``spec = (N, Ctot, blocks)``; a block is
    ``(tensor_index or None, src_sample, samples, dst_sample, dst_sample_stride, dst_channel, channels, fill)``:
    dst[dst_sample + b*stride, dst_channel : dst_channel+channels] <- tensors[i][src_sample + b] (or the constant
    ``fill`` when the index is None).  The blocks must tile dst completely and use every sample of every tensor
    exactly once (checked): the adjoint then overwrites each source gradient in the same single launch."""

    @staticmethod
    def forward(ctx, spec, *tensors):
        N, Ctot, blocks = spec
        dev = tensors[0].device
        assert gather_concat_ok(*tensors), "gather_concat: float32 CUDA NCHW tensors, C*H*W a multiple of 4"
        H, W = tensors[0].shape[-2:]
        covered = [0] * len(tensors)
        dst_elems = 0
        arr = (_lib.CatBlock * len(blocks))()
        for k, (ti, ss, n, ds, dstride, dc, ch, fill) in enumerate(blocks):
            if ti is not None:
                t = tensors[ti]
                assert t.shape[1] == ch and tuple(t.shape[-2:]) == (H, W) and ss + n <= t.shape[0]
                covered[ti] += n
            arr[k] = _lib.CatBlock(_ptr(tensors[ti]) if ti is not None else None, ss,
                                   tensors[ti].stride(0) if ti is not None else 0, ds, dstride, dc, ch, n, float(fill))
            dst_elems += n * ch
        assert dst_elems == N * Ctot, "gather_concat: the blocks do not tile the destination"
        assert all(c == t.shape[0] for c, t in zip(covered, tensors)), "gather_concat: a source sample is unused or used twice"
        with torch.cuda.device(dev):
            dst = torch.empty(N, Ctot, H, W, device=dev, dtype=tensors[0].dtype)
            _lib.call("gather_concat_forward_b200", ctypes.addressof(arr), len(blocks), _ptr(dst), Ctot, H, W, _stream())
        ctx.spec = spec
        ctx.shapes = [tuple(t.shape) for t in tensors]
        return dst

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gdst):
        N, Ctot, blocks = ctx.spec
        gdst = gdst.contiguous()
        H, W = gdst.shape[-2:]
        with torch.cuda.device(gdst.device):
            grads = [torch.empty(sh, device=gdst.device, dtype=gdst.dtype) if need else None
                     for sh, need in zip(ctx.shapes, ctx.needs_input_grad[1:])]
            live = [b for b in blocks if b[0] is not None and grads[b[0]] is not None]
            if live:
                arr = (_lib.CatBlock * len(live))()
                for k, (ti, ss, n, ds, dstride, dc, ch, fill) in enumerate(live):
                    arr[k] = _lib.CatBlock(_ptr(grads[ti]), ss, 0, ds, dstride, dc, ch, n, 0.0)
                _lib.call("gather_concat_backward_b200", ctypes.addressof(arr), len(live), _ptr(gdst), Ctot, H, W, _stream())
        return (None,) + tuple(grads)


def cat_channels(tensors):
    """torch.cat(tensors, dim=1) for contiguous NCHW CUDA tensors through the gather kernel (one launch each way)."""
    N = tensors[0].shape[0]
    blocks, c0 = [], 0
    for i, t in enumerate(tensors):
        blocks.append((i, 0, N, 0, 1, c0, t.shape[1], 0.0))
        c0 += t.shape[1]
    return GatherConcatFunction.apply((N, c0, tuple(blocks)), *tensors)


def gather_concat_ok(*tensors):
    """The gather kernel moves 16-byte vectors out of NCHW tensors whose samples are dense (channel slices of a
    larger tensor are fine: only the sample pitch differs): every sample run (C*H*W floats), every sample pitch and
    every base address must be a multiple of 16 bytes."""
    for t in tensors:
        if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 4):
            return False
        C, H, W = t.shape[1:]
        if (C * H * W) % 4 or t.stride()[1:] != (H * W, W, 1) or t.stride(0) % 4 or t.stride(0) < C * H * W or t.data_ptr() % 16:
            return False
    return True


# ------------------------------------------------------------------------------------------------
# motion-stream prologue (tai.py:67-74; mcnet.py:439-447; util.py:22-41)
# ------------------------------------------------------------------------------------------------

def gray_difference_frames(frames, reverse=False):
    """[B,K,C,H,W] in [-1,1] -> [B,K-1,1,H,W]: gray frames in [0,1], temporal differences (time-reversed input
    order when `reverse`).  One kernel; forward only (the frames are data)."""
    dev = _check("gray_difference_frames", frames)
    B, K, C, H, W = frames.shape
    with torch.cuda.device(dev):
        out = torch.empty(B, K - 1, 1, H, W, device=dev, dtype=frames.dtype)
        _lib.call("gray_difference_frames_b200", _ptr(frames), _ptr(out), B, K, C, H, W, int(bool(reverse)), _stream())
    return out


class GrayDiffPairFunction(torch.autograd.Function):
    """``apply(a, b)`` -> gray01(a) - gray01(b) for [N,C,H,W] frames (mcnet.py:439-447), one kernel each way."""

    @staticmethod
    def forward(ctx, a, b):
        dev = _check("gray_difference_pair", a, b)
        N, C, H, W = a.shape
        assert b.shape == a.shape
        ctx.shape = (N, C, H, W)
        with torch.cuda.device(dev):
            out = torch.empty(N, 1, H, W, device=dev, dtype=a.dtype)
            _lib.call("gray_difference_pair_forward_b200", _ptr(a), _ptr(b), _ptr(out), N, C, H, W, _stream())
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        N, C, H, W = ctx.shape
        grad_out = grad_out.contiguous()
        need_a, need_b = ctx.needs_input_grad
        with torch.cuda.device(grad_out.device):
            ga = torch.empty(N, C, H, W, device=grad_out.device, dtype=grad_out.dtype) if need_a else None
            gb = torch.empty(N, C, H, W, device=grad_out.device, dtype=grad_out.dtype) if need_b else None
            if need_a or need_b:
                _lib.call("gray_difference_pair_backward_b200", _ptr(grad_out), _ptr(ga), _ptr(gb), N, C, H, W, _stream())
        return ga, gb


def frames_to_uint8(frames, flip_channels=None):
    """[..., C, H, W] float in [-1, 1] -> [..., H, W, C] uint8 as predict.py:124-134 forms it (clamp, inverse
    transform, *255, truncation; BGR -> RGB when C == 3 unless flip_channels says otherwise)."""
    dev = _check("frames_to_uint8", frames)
    C, H, W = frames.shape[-3:]
    N = frames.numel() // (C * H * W)
    flip = (C == 3) if flip_channels is None else bool(flip_channels)
    with torch.cuda.device(dev):
        out = torch.empty(frames.shape[:-3] + (H, W, C), device=dev, dtype=torch.uint8)
        _lib.call("frames_to_uint8_b200", _ptr(frames), _ptr(out), N, C, H, W, int(flip), _stream())
    return out


def ffma_probe(grid, block, iters, packed=False):
    """Launch the pure-FFMA probe kernel; returns the sink tensor (flops = 2*8*iters*grid*block)."""
    sink = torch.empty(grid * block, device="cuda", dtype=torch.float32)
    _lib.call("tai_b200_ffma_probe", _ptr(sink), grid, block, iters, int(bool(packed)), _stream())
    return sink
