"""Runs the time-batched Super-SloMo stage kernels (forward and adjoint) at BASELINE config D's shape
([8,3,256,320], T = 3) a few times with cold L2: target for an ncu capture and a quick CUDA-event timing."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_frame_inpainting_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, C, H, W, T = 8, 3, 256, 320, 3
g = torch.Generator(device=dev).manual_seed(0)
R = lambda *s: torch.rand(*s, device=dev, generator=g)
N = lambda *s: torch.randn(*s, device=dev, generator=g)
i0, i1 = R(B, C, H, W), R(B, C, H, W)
f01, f10 = torch.tanh(N(B, 2, H, W)).requires_grad_(True), torch.tanh(N(B, 2, H, W)).requires_grad_(True)
d0, d1 = torch.tanh(N(T * B, 2, H, W)).requires_grad_(True), torch.tanh(N(T * B, 2, H, W)).requires_grad_(True)
v0 = (R(T * B, 1, H, W) * 0.9 + 0.05).requires_grad_(True)
flush = torch.empty(64 * 1024 * 1024, device=dev)
times = {}
for it in range(4):
    flush.zero_()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    ev[0].record()
    X, c0, c1 = ops.SlomoInterpInputFunction.apply(i0, i1, f01, f10, T)
    ev[1].record()
    pred = ops.SlomoRefineBlendFunction.apply(i0, i1, c0, c1, d0, d1, v0, T)
    ev[2].record()
    gp = torch.ones_like(pred)
    gx = torch.ones_like(X)
    flush.zero_()
    ev[4].record()
    grads = torch.autograd.grad([pred, X], [d0, d1, v0, f01, f10], [gp, gx])
    ev[5].record()
    torch.cuda.synchronize()
    if it:
        times.setdefault("interp_input_us", []).append(ev[0].elapsed_time(ev[1]) * 1e3)
        times.setdefault("refine_blend_us", []).append(ev[1].elapsed_time(ev[2]) * 1e3)
        times.setdefault("both_backward_us", []).append(ev[4].elapsed_time(ev[5]) * 1e3)
print(json.dumps({k: min(v) for k, v in times.items()}))
