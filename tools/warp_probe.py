"""Runs the Super-SloMo warp / blend kernels and the bias + activation epilogue a few times at their workload shapes
(target for an ncu capture)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_frame_inpainting_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, C, H, W = 8, 3, 256, 320
g = torch.Generator(device=dev).manual_seed(0)
i0, i1 = torch.rand(B, C, H, W, device=dev, generator=g), torch.rand(B, C, H, W, device=dev, generator=g)
f01, f10 = torch.randn(B, 2, H, W, device=dev, generator=g) * 2, torch.randn(B, 2, H, W, device=dev, generator=g) * 2
d0, d1 = torch.tanh(torch.randn(B, 2, H, W, device=dev, generator=g)), torch.tanh(torch.randn(B, 2, H, W, device=dev, generator=g))
v0 = torch.rand(B, 1, H, W, device=dev, generator=g) * 0.9 + 0.05
go = torch.randn(B, C, H, W, device=dev, generator=g)
y, gy, bias = torch.randn(64, 64, 128, 128, device=dev), torch.randn(64, 64, 128, 128, device=dev), torch.randn(64, device=dev)
flush = torch.empty(64 * 1024 * 1024, device=dev)
for _ in range(3):
    flush.zero_()
    ft0, ft1, g0, g1 = ops.slomo_flow_combine_warp(i0, i1, f01, f10, 0.25)
    ops.slomo_refine_blend(i0, i1, ft0, ft1, d0, d1, v0, 0.25)
    ops.flow_warp_forward(i0, f01)
    ops.flow_warp_backward(i0, f01, go)
    ops.bias_act_forward_(y, bias, "relu", 0.0)
    ops.bias_act_backward(gy, y, "relu", 0.0)
torch.cuda.synchronize()
