"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the TAI / bi-TAI hot path.

A NumPy (float64) restatement of the arithmetic of the reference
(MichiganCOG/video-frame-inpainting) for the path named in BASELINE.json, plus a ctypes
front-end to the C restatement in ``oracle/sepconv_oracle.c``.  Each function cites the
reference ``file:line`` it follows.

Who may import this module: ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` -- always as the checker (or the timed CPU baseline),
never as part of the product path in ``video_frame_inpainting_b200``.

Parity status.  The reference has no tests, golden vectors or CPU implementation for this path
 and cannot be imported under Python 3 / torch 2.
The oracle is therefore pinned as follows:
  * separable convolution (forward, gradV, gradH, gradI): pinned against the reference's own
    UNMODIFIED CUDA kernels, compiled from /root/reference into ``oracle/_ref`` and run on a
    B200 (tests/test_ref_kernels_gpu.py; fixtures from that run are in tests/golden/).
  * replication pad, blend, ConvLSTM gates, Super-SloMo warp/blend: the reference expresses
    these with torch-0.3.1 library ops (third-party, un-vendored: torch==0.3.1,
    requirements.txt:25).  They are restated from the published semantics of those ops and
    cross-checked against torch 2.x CPU ops where the semantics are unchanged
    (tests/test_oracle_cpu.py).  For ``grid_sample`` the 0.3.1 mapping (bilinear, zero padding,
    ``ix = ((g+1)/2)*(W-1)``) cannot be verified offline: **parity unpinned** for the warp beyond
    agreement with ``F.grid_sample(..., align_corners=True)`` of torch 2.x.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


# --------------------------------------------------------------------------------------------
# C restatement (oracle/sepconv_oracle.c)
# --------------------------------------------------------------------------------------------

def build_c_oracle(force: bool = False) -> str:
    """Compile oracle/sepconv_oracle.c with gcc (OpenMP if available).  Returns the .so path."""
    src = os.path.join(_HERE, "sepconv_oracle.c")
    if (not force and os.path.isfile(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    os.makedirs(os.path.dirname(_LIB_PATH), exist_ok=True)
    base = ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c99", src, "-o", _LIB_PATH]
    try:
        subprocess.run(base[:1] + ["-fopenmp"] + base[1:], check=True, capture_output=True)
    except (subprocess.CalledProcessError, FileNotFoundError):
        subprocess.run(base, check=True, capture_output=True)
    return _LIB_PATH


def c_lib():
    global _lib
    if _lib is None:
        build_c_oracle()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _fptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f32c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _dims(inp, ver, hor, ks):
    B, C, Hi, Wi = inp.shape
    Ho, Wo = Hi - ks + 1, Wi - ks + 1
    assert ver.shape == (B, ks, Ho, Wo) and hor.shape == (B, ks, Ho, Wo), (inp.shape, ver.shape, hor.shape, ks)
    return B, C, Hi, Wi, Ho, Wo


def sepconv_forward(inp, ver, hor, ks, dtype=np.float64):
    """O[b,c,y,x] = sum_i sum_j I[b,c,y+i,x+j] V[b,i,y,x] H[b,j,y,x]   (kernel.cu:19-47).

    dtype=float64 -> the oracle; dtype=float32 -> the literal FP32 port (CPU baseline)."""
    inp, ver, hor = _f32c(inp), _f32c(ver), _f32c(hor)
    B, C, Hi, Wi, Ho, Wo = _dims(inp, ver, hor, ks)
    out = np.empty((B, C, Ho, Wo), dtype=dtype)
    fn = c_lib().oracle_sepconv_forward_f64 if dtype == np.float64 else c_lib().oracle_sepconv_forward_f32
    fn(_fptr(inp), _fptr(ver), _fptr(hor), _fptr(out), B, C, Hi, Wi, ks)
    return out


def sepconv_grad_vertical(gout, inp, hor, ks, dtype=np.float64):
    """gV[b,i,y,x] = sum_c sum_j gO[b,c,y,x] I[b,c,y+i,x+j] H[b,j,y,x]   (kernel.cu:49-86)."""
    gout, inp, hor = _f32c(gout), _f32c(inp), _f32c(hor)
    B, C, Hi, Wi = inp.shape
    Ho, Wo = Hi - ks + 1, Wi - ks + 1
    assert gout.shape == (B, C, Ho, Wo) and hor.shape == (B, ks, Ho, Wo)
    g = np.empty((B, ks, Ho, Wo), dtype=dtype)
    fn = (c_lib().oracle_sepconv_grad_vertical_f64 if dtype == np.float64
          else c_lib().oracle_sepconv_grad_vertical_f32)
    fn(_fptr(gout), _fptr(inp), _fptr(hor), _fptr(g), B, C, Hi, Wi, ks)
    return g


def sepconv_grad_horizontal(gout, inp, ver, ks, dtype=np.float64):
    """gH[b,j,y,x] = sum_c sum_i gO[b,c,y,x] I[b,c,y+i,x+j] V[b,i,y,x]   (kernel.cu:88-118)."""
    gout, inp, ver = _f32c(gout), _f32c(inp), _f32c(ver)
    B, C, Hi, Wi = inp.shape
    Ho, Wo = Hi - ks + 1, Wi - ks + 1
    assert gout.shape == (B, C, Ho, Wo) and ver.shape == (B, ks, Ho, Wo)
    g = np.empty((B, ks, Ho, Wo), dtype=dtype)
    fn = (c_lib().oracle_sepconv_grad_horizontal_f64 if dtype == np.float64
          else c_lib().oracle_sepconv_grad_horizontal_f32)
    fn(_fptr(gout), _fptr(inp), _fptr(ver), _fptr(g), B, C, Hi, Wi, ks)
    return g


def sepconv_grad_input(gout, ver, hor, ks, dtype=np.float64):
    """gI over the PADDED input, with the bounds test of kernel.cu:150   (kernel.cu:120-162)."""
    gout, ver, hor = _f32c(gout), _f32c(ver), _f32c(hor)
    B, C, Ho, Wo = gout.shape
    Hi, Wi = Ho + ks - 1, Wo + ks - 1
    assert ver.shape == (B, ks, Ho, Wo) and hor.shape == (B, ks, Ho, Wo)
    g = np.empty((B, C, Hi, Wi), dtype=dtype)
    fn = (c_lib().oracle_sepconv_grad_input_f64 if dtype == np.float64
          else c_lib().oracle_sepconv_grad_input_f32)
    fn(_fptr(gout), _fptr(ver), _fptr(hor), _fptr(g), B, C, Hi, Wi, ks)
    return g


def sepconv_grad_input_tapcount(Hi, Wi, ks):
    """Integer table: taps passing the bounds test of kernel.cu:150, per padded-input element."""
    cnt = np.empty((Hi, Wi), dtype=np.int32)
    c_lib().oracle_sepconv_grad_input_tapcount(_fptr(cnt), Hi, Wi, ks)
    return cnt


def replication_pad_index(H, W, p):
    """Integer tables (src_y[H+2p], src_x[W+2p]) of torch.nn.ReplicationPad2d(p)  (tai.py:170-171)."""
    sy = np.empty(H + 2 * p, dtype=np.int32)
    sx = np.empty(W + 2 * p, dtype=np.int32)
    c_lib().oracle_replication_pad_index(_fptr(sy), _fptr(sx), H, W, p)
    return sy, sx


# --------------------------------------------------------------------------------------------
# Independent NumPy restatement of the separable convolution (small cases; cross-checks the C)
# --------------------------------------------------------------------------------------------

def sepconv_forward_np(inp, ver, hor, ks):
    """Same math as kernel.cu:19-47, vectorised with a sliding window (float64)."""
    inp = np.asarray(inp, np.float64)
    ver = np.asarray(ver, np.float64)
    hor = np.asarray(hor, np.float64)
    win = np.lib.stride_tricks.sliding_window_view(inp, (ks, ks), axis=(2, 3))  # B,C,Ho,Wo,i,j
    return np.einsum("bcyxij,biyx,bjyx->bcyx", win, ver, hor, optimize=True)


def sepconv_backward_np(gout, inp, ver, hor, ks):
    """(gI, gV, gH) as in kernel.cu:49-162, float64, written as adjoints of sepconv_forward_np."""
    gout = np.asarray(gout, np.float64)
    inp = np.asarray(inp, np.float64)
    ver = np.asarray(ver, np.float64)
    hor = np.asarray(hor, np.float64)
    B, C, Hi, Wi = inp.shape
    Ho, Wo = Hi - ks + 1, Wi - ks + 1
    win = np.lib.stride_tricks.sliding_window_view(inp, (ks, ks), axis=(2, 3))
    gver = np.einsum("bcyx,bcyxij,bjyx->biyx", gout, win, hor, optimize=True)
    ghor = np.einsum("bcyx,bcyxij,biyx->bjyx", gout, win, ver, optimize=True)
    gin = np.zeros((B, C, Hi, Wi), np.float64)
    for i in range(ks):
        for j in range(ks):
            gin[:, :, i:i + Ho, j:j + Wo] += gout * (ver[:, i] * hor[:, j])[:, None]
    return gin, gver, ghor


# --------------------------------------------------------------------------------------------
# Replication pad (tai.py:170-171,229,233) and its adjoint
# --------------------------------------------------------------------------------------------

def replication_pad(img, p):
    """I_pad[y,x] = I[clamp(y-p,0,H-1), clamp(x-p,0,W-1)] -- torch.nn.ReplicationPad2d(p)."""
    H, W = img.shape[-2:]
    sy, sx = replication_pad_index(H, W, p)
    return img[..., sy[:, None], sx[None, :]]


def replication_pad_adjoint(gpad, p):
    """Backward of ReplicationPad2d: scatter-add every padded element onto its clamped source."""
    Hp, Wp = gpad.shape[-2:]
    H, W = Hp - 2 * p, Wp - 2 * p
    sy, sx = replication_pad_index(H, W, p)
    rows = np.zeros(gpad.shape[:-2] + (H, Wp), dtype=np.float64)
    np.add.at(rows, (Ellipsis, sy, slice(None)), np.asarray(gpad, np.float64))
    out = np.zeros(gpad.shape[:-2] + (H, W), dtype=np.float64)
    np.add.at(out, (Ellipsis, sx), rows)
    return out


# --------------------------------------------------------------------------------------------
# TAI / TWI blend (tai.py:90,99,105; twi.py:90,105)
# --------------------------------------------------------------------------------------------

def time_weights(T):
    """w = np.linspace(0, 1, num=T+2).tolist()[1:-1]   (tai.py:90, twi.py:90)."""
    return np.linspace(0, 1, num=T + 2).tolist()[1:-1]


def tai_blend(dot1, dot2, a=0.5, b=0.5):
    """pred = a*Dot1 + b*Dot2; TAI: a=b=0.5 (tai.py:105); TWI: a=1-w_t, b=w_t (twi.py:105)."""
    return a * np.asarray(dot1, np.float64) + b * np.asarray(dot2, np.float64)


def tai_fused_forward(pred_f, pred_b, v1, h1, v2, h2, ks, a=0.5, b=0.5):
    """pad -> sepconv (both streams) -> blend, i.e. tai.py:229-236 followed by tai.py:105."""
    p = ks // 2
    d1 = sepconv_forward(replication_pad(np.asarray(pred_f, np.float32), p), v1, h1, ks)
    d2 = sepconv_forward(replication_pad(np.asarray(pred_b, np.float32), p), v2, h2, ks)
    return tai_blend(d1, d2, a, b), d1, d2


# --------------------------------------------------------------------------------------------
# ConvLSTM gates (mcnet.py:281-294)
# --------------------------------------------------------------------------------------------

def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def convlstm_gates(conv_out, state, forget_bias=1.0):
    """c,h = chunk(state,2); (i,j,f,o) = chunk(conv_out,4);
    c' = c*sigmoid(f+forget_bias) + sigmoid(i)*tanh(j);  h' = tanh(c')*sigmoid(o);
    returns (h', cat(c',h'))                                          (mcnet.py:287-294)."""
    conv_out = np.asarray(conv_out, np.float64)
    state = np.asarray(state, np.float64)
    F = state.shape[1] // 2
    c = state[:, :F]
    i, j, f, o = (conv_out[:, k * F:(k + 1) * F] for k in range(4))
    new_c = c * _sigmoid(f + forget_bias) + _sigmoid(i) * np.tanh(j)
    new_h = np.tanh(new_c) * _sigmoid(o)
    return new_h, np.concatenate([new_c, new_h], axis=1)


def convlstm_gates_backward(conv_out, state, g_new_state, forget_bias=1.0):
    """Adjoint of convlstm_gates w.r.t. (conv_out, state), given d/d new_state = cat(gc', gh')."""
    conv_out = np.asarray(conv_out, np.float64)
    state = np.asarray(state, np.float64)
    g = np.asarray(g_new_state, np.float64)
    F = state.shape[1] // 2
    c = state[:, :F]
    i, j, f, o = (conv_out[:, k * F:(k + 1) * F] for k in range(4))
    si, sf, so, tj = _sigmoid(i), _sigmoid(f + forget_bias), _sigmoid(o), np.tanh(j)
    new_c = c * sf + si * tj
    tc = np.tanh(new_c)
    gc_new, gh_new = g[:, :F], g[:, F:]
    gct = gc_new + gh_new * so * (1.0 - tc * tc)
    g_conv = np.concatenate([
        gct * tj * si * (1.0 - si),
        gct * si * (1.0 - tj * tj),
        gct * c * sf * (1.0 - sf),
        gh_new * tc * so * (1.0 - so)], axis=1)
    g_state = np.concatenate([gct * sf, np.zeros_like(gh_new)], axis=1)
    return g_conv, g_state


# --------------------------------------------------------------------------------------------
# Super SloMo: bilinear backward warp (slomo.py:265-286) through torch-0.3.1 grid_sample
# --------------------------------------------------------------------------------------------

def flow_warp_coords_f32(uv):
    """The FP32 sampling coordinates exactly as the reference computes them, one rounding per
    operation, no contraction:  X = x + u;  g = 2*(X/W - 0.5)              (slomo.py:279-282)
    then grid_sample 0.3.1:     ix = ((g + 1) / 2) * (W - 1)
    Returns (ix, iy, ix0, iy0) with ix0 = floor(ix) as int32 -- the integer table that the CUDA
    kernel must reproduce bit-exactly."""
    uv = np.asarray(uv, np.float32)
    B, _, H, W = uv.shape
    f = np.float32
    gx = np.arange(W, dtype=np.float32)[None, None, :]
    gy = np.arange(H, dtype=np.float32)[None, :, None]
    X = (gx + uv[:, 0]).astype(f)
    Y = (gy + uv[:, 1]).astype(f)
    X = (f(2) * ((X / f(W)).astype(f) - f(0.5)).astype(f)).astype(f)
    Y = (f(2) * ((Y / f(H)).astype(f) - f(0.5)).astype(f)).astype(f)
    ix = ((((X + f(1)).astype(f)) / f(2)).astype(f) * f(W - 1)).astype(f)
    iy = ((((Y + f(1)).astype(f)) / f(2)).astype(f) * f(H - 1)).astype(f)
    return ix, iy, np.floor(ix).astype(np.int32), np.floor(iy).astype(np.int32)


def _gather_zero(img, yy, xx):
    """img[b,c,yy,xx] with zero padding outside the image (SAFE_GET of the 0.3.1 sampler)."""
    B, C, H, W = img.shape
    ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
    yc = np.clip(yy, 0, H - 1)
    xc = np.clip(xx, 0, W - 1)
    bidx = np.arange(B)[:, None, None]
    vals = img[bidx, :, yc, xc]                       # B,H,W,C
    vals = np.where(ok[..., None], vals, 0.0)
    return np.moveaxis(vals, -1, 1), ok


def flow_warp(img, uv, coords="f64"):
    """out[b,c,y,x] = bilinear(img[b,c], ix, iy), zero padding             (slomo.py:265-286).

    coords="f64": coordinates in float64 (value oracle).  coords="f32": coordinates and corner
    indices from flow_warp_coords_f32 (integer-exact twin of the CUDA kernel), weights and
    accumulation still in float64."""
    img = np.asarray(img, np.float64)
    B, C, H, W = img.shape
    if coords == "f32":
        ix, iy, x0, y0 = flow_warp_coords_f32(uv)
        ix, iy = ix.astype(np.float64), iy.astype(np.float64)
    else:
        uv = np.asarray(uv, np.float64)
        X = np.arange(W, dtype=np.float64)[None, None, :] + uv[:, 0]
        Y = np.arange(H, dtype=np.float64)[None, :, None] + uv[:, 1]
        ix = ((2 * (X / W - 0.5) + 1) / 2) * (W - 1)
        iy = ((2 * (Y / H - 0.5) + 1) / 2) * (H - 1)
        x0, y0 = np.floor(ix).astype(np.int64), np.floor(iy).astype(np.int64)
    x1, y1 = x0 + 1, y0 + 1
    w_nw = (x1 - ix) * (y1 - iy)
    w_ne = (ix - x0) * (y1 - iy)
    w_sw = (x1 - ix) * (iy - y0)
    w_se = (ix - x0) * (iy - y0)
    nw, _ = _gather_zero(img, y0, x0)
    ne, _ = _gather_zero(img, y0, x1)
    sw, _ = _gather_zero(img, y1, x0)
    se, _ = _gather_zero(img, y1, x1)
    return nw * w_nw[:, None] + ne * w_ne[:, None] + sw * w_sw[:, None] + se * w_se[:, None]


def flow_warp_backward(img, uv, gout):
    """Adjoint of flow_warp w.r.t. (img, uv); float64.  d ix / d u = (W-1)/W, d iy / d v = (H-1)/H."""
    img = np.asarray(img, np.float64)
    uv = np.asarray(uv, np.float64)
    gout = np.asarray(gout, np.float64)
    B, C, H, W = img.shape
    X = np.arange(W, dtype=np.float64)[None, None, :] + uv[:, 0]
    Y = np.arange(H, dtype=np.float64)[None, :, None] + uv[:, 1]
    ix = ((2 * (X / W - 0.5) + 1) / 2) * (W - 1)
    iy = ((2 * (Y / H - 0.5) + 1) / 2) * (H - 1)
    x0, y0 = np.floor(ix).astype(np.int64), np.floor(iy).astype(np.int64)
    x1, y1 = x0 + 1, y0 + 1
    gimg = np.zeros_like(img)
    guv = np.zeros_like(uv)
    bidx = np.broadcast_to(np.arange(B)[:, None, None], x0.shape)
    for (yy, xx, wgt, dwx, dwy) in (
            (y0, x0, (x1 - ix) * (y1 - iy), -(y1 - iy), -(x1 - ix)),
            (y0, x1, (ix - x0) * (y1 - iy), (y1 - iy), -(ix - x0)),
            (y1, x0, (x1 - ix) * (iy - y0), -(iy - y0), (x1 - ix)),
            (y1, x1, (ix - x0) * (iy - y0), (iy - y0), (ix - x0))):
        vals, ok = _gather_zero(img, yy, xx)
        for c in range(C):
            contrib = np.where(ok, gout[:, c] * wgt, 0.0)
            np.add.at(gimg[:, c], (bidx, np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)), contrib)
        guv[:, 0] += (vals * gout).sum(1) * dwx * (W - 1) / W
        guv[:, 1] += (vals * gout).sum(1) * dwy * (H - 1) / H
    return gimg, guv


def slomo_flow_combine(f01, f10, t):
    """F_t0 = -(1-t) t F01 + t^2 F10;  F_t1 = (1-t)^2 F01 - t (1-t) F10     (slomo.py:312-314),
    with t = (t_ + 1) / (T + 1) in true division (slomo.py:2,312)."""
    f01 = np.asarray(f01, np.float64)
    f10 = np.asarray(f10, np.float64)
    return (-(1 - t) * t * f01 + t ** 2 * f10,
            (1 - t) * (1 - t) * f01 - t * (1 - t) * f10)


def slomo_refine_blend(i0, i1, f_t0, f_t1, d_t0, d_t1, v_t0, t):
    """F_ref = clamp(dF + F, -1, 1); V1 = 1 - V0; g = warp(I, F_ref);
    out = ((1-t) V0 g0 + t V1 g1) / ((1-t) V0 + t V1)                       (slomo.py:320-328)."""
    r0 = np.clip(np.asarray(d_t0, np.float64) + np.asarray(f_t0, np.float64), -1, 1)
    r1 = np.clip(np.asarray(d_t1, np.float64) + np.asarray(f_t1, np.float64), -1, 1)
    v0 = np.asarray(v_t0, np.float64)
    v1 = 1 - v0
    g0 = flow_warp(i0, r0)
    g1 = flow_warp(i1, r1)
    norm = (1 - t) * v0 + t * v1
    return ((1 - t) * v0 * g0 + t * v1 * g1) / norm


def slomo_interp_input(i0, i1, f01, f10, T):
    """The per-t loop of slomo.py:307-318 for all T middle frames: returns
    X [T*B, 4C+4, H, W] = cat(I0, g(I0,F_t0), F_t0, F_t1, g(I1,F_t1), I1) per t (sample n = t_*B + b), and the
    collectors F_t_0 / F_t_1 [B,T,2,H,W] in the reference's REVERSED time order (new frames are prepended,
    slomo.py:332-340: slot T-1-t_)."""
    i0, i1 = np.asarray(i0, np.float64), np.asarray(i1, np.float64)
    B, C, H, W = i0.shape
    X = np.zeros((T * B, 4 * C + 4, H, W))
    c0, c1 = np.zeros((B, T, 2, H, W)), np.zeros((B, T, 2, H, W))
    for t_ in range(T):
        t = (t_ + 1) / (T + 1)
        ft0, ft1 = slomo_flow_combine(f01, f10, t)
        X[t_ * B:(t_ + 1) * B] = np.concatenate((i0, flow_warp(i0, ft0), ft0, ft1, flow_warp(i1, ft1), i1), 1)
        c0[:, T - 1 - t_], c1[:, T - 1 - t_] = ft0, ft1
    return X, c0, c1


def slomo_interp_input_backward(i0, i1, f01, f10, T, gX, gc0=None, gc1=None):
    """Adjoint of slomo_interp_input w.r.t. (F_0_1, F_1_0); float64."""
    i0, i1 = np.asarray(i0, np.float64), np.asarray(i1, np.float64)
    f01, f10, gX = np.asarray(f01, np.float64), np.asarray(f10, np.float64), np.asarray(gX, np.float64)
    B, C, H, W = i0.shape
    g01, g10 = np.zeros_like(f01), np.zeros_like(f10)
    for t_ in range(T):
        t = (t_ + 1) / (T + 1)
        ft0, ft1 = slomo_flow_combine(f01, f10, t)
        g = gX[t_ * B:(t_ + 1) * B]
        g0 = flow_warp_backward(i0, ft0, g[:, C:2 * C])[1] + g[:, 2 * C:2 * C + 2]
        g1 = flow_warp_backward(i1, ft1, g[:, 2 * C + 4:3 * C + 4])[1] + g[:, 2 * C + 2:2 * C + 4]
        if gc0 is not None:
            g0 = g0 + np.asarray(gc0, np.float64)[:, T - 1 - t_]
        if gc1 is not None:
            g1 = g1 + np.asarray(gc1, np.float64)[:, T - 1 - t_]
        g01 += -(1 - t) * t * g0 + (1 - t) * (1 - t) * g1
        g10 += t ** 2 * g0 - t * (1 - t) * g1
    return g01, g10


def slomo_refine_blend_batched(i0, i1, c0, c1, d0, d1, v0, T):
    """slomo.py:320-340 for all T middle frames: pred [B,T,C,H,W] in reversed time order; the refinement outputs
    d0, d1 [T*B,2,H,W], v0 [T*B,1,H,W] in sample order n = t_*B + b, the flows from the collectors."""
    B, C, H, W = np.asarray(i0).shape
    pred = np.zeros((B, T, C, H, W))
    for t_ in range(T):
        t = (t_ + 1) / (T + 1)
        sl = slice(t_ * B, (t_ + 1) * B)
        pred[:, T - 1 - t_] = slomo_refine_blend(i0, i1, np.asarray(c0)[:, T - 1 - t_], np.asarray(c1)[:, T - 1 - t_],
                                                 np.asarray(d0)[sl], np.asarray(d1)[sl], np.asarray(v0)[sl], t)
    return pred


def slomo_refine_blend_batched_backward(i0, i1, c0, c1, d0, d1, v0, T, gpred):
    """Adjoint of slomo_refine_blend_batched w.r.t. (collectors, d0, d1, v0); float64.  torch.clamp passes the
    gradient on [-1, 1] inclusive."""
    i0, i1 = np.asarray(i0, np.float64), np.asarray(i1, np.float64)
    c0, c1, d0, d1, v0, gpred = (np.asarray(a, np.float64) for a in (c0, c1, d0, d1, v0, gpred))
    B, C, H, W = i0.shape
    gc0, gc1, gd0, gd1, gv0 = (np.zeros_like(a) for a in (c0, c1, d0, d1, v0))
    for t_ in range(T):
        t = (t_ + 1) / (T + 1)
        sl, slot = slice(t_ * B, (t_ + 1) * B), T - 1 - t_
        s0, s1 = d0[sl] + c0[:, slot], d1[sl] + c1[:, slot]
        r0, r1 = np.clip(s0, -1, 1), np.clip(s1, -1, 1)
        a0, a1 = flow_warp(i0, r0), flow_warp(i1, r1)
        V = v0[sl]
        k0, k1 = (1 - t) * V, t * (1 - V)
        n = k0 + k1
        g = gpred[:, slot]
        out = (k0 * a0 + k1 * a1) / n
        gv0[sl] = (g * (((1 - t) * a0 - t * a1) - out * ((1 - t) - t)) / n).sum(1, keepdims=True)
        gr0 = flow_warp_backward(i0, r0, g * k0 / n)[1] * ((s0 >= -1) & (s0 <= 1))
        gr1 = flow_warp_backward(i1, r1, g * k1 / n)[1] * ((s1 >= -1) & (s1 <= 1))
        gd0[sl], gd1[sl] = gr0, gr1
        gc0[:, slot], gc1[:, slot] = gr0, gr1
    return gc0, gc1, gd0, gd1, gv0


# --------------------------------------------------------------------------------------------
# Small glue on the call path (util.py:22-41; mcnet.py:240-256)
# --------------------------------------------------------------------------------------------

def inverse_transform(x):
    """(x + 1) / 2   (util.py:22-23)."""
    return (np.asarray(x, np.float64) + 1.0) / 2


def bgr2gray(x, axis=1):
    """0.1140 B + 0.5870 G + 0.2989 R over the channel axis, keepdim   (util.py:30-41)."""
    x = np.asarray(x, np.float64)
    b, g, r = (np.take(x, k, axis=axis) for k in range(3))
    return np.expand_dims(0.1140 * b + 0.5870 * g + 0.2989 * r, axis)


def fixed_unpooling(x):
    """out[2y,2x] = x[y,x], the other three of every 2x2 cell are zero   (mcnet.py:240-256)."""
    x = np.asarray(x, np.float64)
    B, C, H, W = x.shape
    out = np.zeros((B, C, 2 * H, 2 * W), np.float64)
    out[:, :, ::2, ::2] = x
    return out


def maxpool2x2(x):
    """nn.MaxPool2d(2) (mcnet.py:28-45; slomo.py:47-85): floor mode.  Returns (values, code) with code =
    2*dy+dx of the selected element under the library kernel's rule: scan (0,0),(0,1),(1,0),(1,1), a later
    element replaces the current one only if it is greater or NaN (first maximum wins ties)."""
    x = np.asarray(x)
    H, W = x.shape[-2:]
    Ho, Wo = H // 2, W // 2
    win = [x[..., dy:2 * Ho:2, dx:2 * Wo:2] for dy in (0, 1) for dx in (0, 1)]
    m = win[0].copy()
    code = np.zeros(m.shape, np.uint8)
    for k in (1, 2, 3):
        take = (win[k] > m) | np.isnan(win[k])
        m = np.where(take, win[k], m)
        code = np.where(take, np.uint8(k), code)
    return m, code


def maxpool2x2_backward(gout, code, H, W):
    """Gradient routed to the selected element; zeros elsewhere (and in the unpooled last row / column)."""
    gout = np.asarray(gout)
    g = np.zeros(gout.shape[:-2] + (H, W), gout.dtype)
    Ho, Wo = H // 2, W // 2
    for k in range(4):
        dy, dx = k // 2, k % 2
        g[..., dy:2 * Ho:2, dx:2 * Wo:2] = np.where(code == k, gout, 0)
    return g


def upsample_bilinear2x_taps(n_in):
    """Integer taps and FP32 weights of the torch-0.3.1 bilinear x2 upsample along one axis (the library
    computed src = d * ((in-1)/(out-1)) in FP32, i0 = (int)src, i1 = i0 + (i0 < in-1), lambda = src - i0).
    Returns (i0, i1, w0, w1) for d = 0 .. 2*n_in-1; the integer arrays are compared bit-exactly."""
    n_out = 2 * n_in
    ratio = np.float32(n_in - 1) / np.float32(n_out - 1) if n_out > 1 else np.float32(0)
    d = np.arange(n_out, dtype=np.float32)
    src = (ratio * d).astype(np.float32)
    i0 = src.astype(np.int64)
    i1 = i0 + (i0 < n_in - 1)
    w1 = (src - i0.astype(np.float32)).astype(np.float32)
    w0 = (np.float32(1) - w1).astype(np.float32)
    return i0, i1, w0.astype(np.float64), w1.astype(np.float64)


def upsample_bilinear2x(x):
    """nn.Upsample(scale_factor=2, mode='bilinear') of torch 0.3.1 == align_corners=True   (tai.py:283,337,343;
    slomo.py:113-149): weights as the library formed them (FP32), products and sums in float64."""
    x = np.asarray(x, np.float64)
    H, W = x.shape[-2:]
    y0, y1, wy0, wy1 = upsample_bilinear2x_taps(H)
    x0, x1, wx0, wx1 = upsample_bilinear2x_taps(W)
    top = x[..., y0, :][..., :, x0] * wx0 + x[..., y0, :][..., :, x1] * wx1
    bot = x[..., y1, :][..., :, x0] * wx0 + x[..., y1, :][..., :, x1] * wx1
    return top * wy0[:, None] + bot * wy1[:, None]


def upsample_bilinear2x_backward(g):
    """Adjoint of upsample_bilinear2x: scatter every output gradient to its four taps (float64)."""
    g = np.asarray(g, np.float64)
    Ho, Wo = g.shape[-2:]
    H, W = Ho // 2, Wo // 2
    y0, y1, wy0, wy1 = upsample_bilinear2x_taps(H)
    x0, x1, wx0, wx1 = upsample_bilinear2x_taps(W)
    rows = np.zeros(g.shape[:-2] + (H, Wo))
    np.add.at(rows, (Ellipsis, y0, slice(None)), g * wy0[:, None])
    np.add.at(rows, (Ellipsis, y1, slice(None)), g * wy1[:, None])
    out = np.zeros(g.shape[:-2] + (H, W))
    rt = np.swapaxes(rows, -1, -2)  # [..., Wo, H]
    ot = np.swapaxes(out, -1, -2)   # view [..., W, H]
    np.add.at(ot, (Ellipsis, x0, slice(None)), rt * wx0[:, None])
    np.add.at(ot, (Ellipsis, x1, slice(None)), rt * wx1[:, None])
    return out


# --------------------------------------------------------------------------------------------
# reconstruction losses of the training step
# --------------------------------------------------------------------------------------------

def _gdl_args(x01, y01):
    """The arguments of the two |.| terms of GDL (losses.py:30-35), evaluated in FP32 with one rounding per
    reference operation so that their SIGNS (the integer part of the backward) are the reference's own."""
    x = np.asarray(x01, np.float32)
    y = np.asarray(y01, np.float32)
    H, W = x.shape[-2:]
    x = x.reshape(-1, H, W)
    y = y.reshape(-1, H, W)
    w_arg = ((x[:, :, :-1] - x[:, :, 1:]) - (y[:, :, :-1] - y[:, :, 1:]))[:, 1:, :]   # rows 1.., cols 0..W-2
    h_arg = ((x[:, 1:, :] - x[:, :-1, :]) - (y[:, 1:, :] - y[:, :-1, :]))[:, :, 1:]   # rows 0..H-2, cols 1..
    return x, y, w_arg, h_arg


def l2_gdl_loss(pred, target, add=1.0, mul=0.5):
    """(MSELoss, GDL) of the inverse-transformed tensors   (environments.py:363-371; util.py:22-23;
    losses.py:24-45 with reduce=True).  Terms in FP32 like the reference, sums in float64."""
    x01 = ((np.asarray(pred, np.float32) + np.float32(add)) * np.float32(mul)).astype(np.float32)
    y01 = ((np.asarray(target, np.float32) + np.float32(add)) * np.float32(mul)).astype(np.float32)
    x, y, w_arg, h_arg = _gdl_args(x01, y01)
    d = (x - y).astype(np.float64)
    mse = float(np.mean(d * d))
    n_gdl = w_arg.size
    gdl = float((np.abs(w_arg).astype(np.float64).sum() + np.abs(h_arg).astype(np.float64).sum()) / n_gdl) if n_gdl else float("nan")
    return mse, gdl


def l2_gdl_loss_backward(pred, target, g_mse, g_gdl, add=1.0, mul=0.5):
    """d(g_mse * mse + g_gdl * gdl) / d pred; sign(0) = 0 as in torch's abs backward."""
    shape = np.asarray(pred).shape
    x01 = ((np.asarray(pred, np.float32) + np.float32(add)) * np.float32(mul)).astype(np.float32)
    y01 = ((np.asarray(target, np.float32) + np.float32(add)) * np.float32(mul)).astype(np.float32)
    x, y, w_arg, h_arg = _gdl_args(x01, y01)
    g = 2.0 * g_mse * mul / x.size * (x - y).astype(np.float64)
    if w_arg.size:
        cg = g_gdl * mul / w_arg.size
        sw = np.sign(w_arg).astype(np.float64)
        sh = np.sign(h_arg).astype(np.float64)
        g[:, 1:, :-1] += cg * sw     # + on the left column of the pair
        g[:, 1:, 1:] -= cg * sw      # - on the right column
        g[:, 1:, 1:] += cg * sh      # + on the lower row of the pair
        g[:, :-1, 1:] -= cg * sh     # - on the upper row
    return g.reshape(shape)


def frames_to_uint8(video):
    """predict.py:124-134 (save_video_frames) up to the PNG encoder: [T,C,H,W] float in [-1,1] ->
    [T,H,W,C] uint8, FP32 arithmetic as numpy evaluates the reference expression, RGB order for C == 3."""
    v = np.clip(np.asarray(video, np.float32), np.float32(-1), np.float32(1))
    frames = np.transpose(v, (0, 2, 3, 1))                          # to_numpy(..., transpose=(1, 2, 0)) per frame
    u8 = (255 * ((frames + 1.) / 2)).astype(np.uint8)               # util.py:22-23 inside predict.py:132
    return u8[..., ::-1] if video.shape[1] == 3 else u8


def rel_err(x, ref):
    """max |x-ref| / max(|ref|, rms(ref)) -- the tolerance definition of SURVEY.md section 7
    (V/H are unnormalised and signed, so outputs have zero crossings)."""
    ref = np.asarray(ref, np.float64)
    x = np.asarray(x, np.float64)
    scale = np.maximum(np.abs(ref), np.sqrt(np.mean(ref * ref)) + 1e-30)
    return float(np.max(np.abs(x - ref) / scale)) if ref.size else 0.0
