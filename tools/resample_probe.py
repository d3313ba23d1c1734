"""Runs the resampling / streaming kernels a few times at the model's largest shapes (target for an ncu capture)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_frame_inpainting_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, C, H, W = 32, 64, 64, 64
x = torch.randn(B, C, H, W, device=dev)
g = torch.randn(B, C, 2 * H, 2 * W, device=dev)
flush = torch.empty(64 * 1024 * 1024, device=dev)
conv, state = torch.randn(32, 1024, 16, 16, device=dev), torch.randn(32, 512, 16, 16, device=dev)
gns = torch.randn_like(state)
a, b = torch.rand(160, 1, 128, 128, device=dev) * 2 - 1, torch.rand(160, 1, 128, 128, device=dev) * 2 - 1
one = torch.ones(1, device=dev)
bias, gy = torch.randn(64, device=dev), torch.randn(64, 64, 128, 128, device=dev)
for _ in range(3):
    flush.zero_()
    ops.upsample_bilinear2x_backward(g)
    ops.upsample_bilinear2x_forward(x)
    ops.convlstm_gates_forward(conv, state, 1.0)
    ops.convlstm_gates_backward(conv, state, gns, 1.0)
    ops.unpool_backward(g)
    ops.unpool_add_forward(x, g)
    out, code = ops.maxpool2x2_forward(g)
    ops.maxpool2x2_backward(out, code, 2 * H, 2 * W)
    ops.l2_gdl_loss_forward(a, b)
    ops.l2_gdl_loss_backward(a, b, one, one)
    y = torch.randn(64, 64, 128, 128, device=dev)
    ops.bias_act_forward_(y, bias, "relu", 0.0)
    ops.bias_act_backward(gy, y, "relu", 0.0)
torch.cuda.synchronize()
