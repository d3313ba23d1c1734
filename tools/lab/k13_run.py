"""lab: forward kernel timing at ks = 13 (HBM-bound class) for a few shapes; run once per library variant."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from video_frame_inpainting_b200 import ops
dev = torch.device("cuda:0")
flush = torch.empty(64 * 1024 * 1024, device=dev)
for (B, C, S) in [(16, 1, 128), (64, 1, 256), (64, 1, 512), (64, 3, 256), (16, 3, 512)]:
    ks = 13
    I = torch.rand(B, C, S + ks - 1, S + ks - 1, device=dev)
    V, H = torch.rand(B, ks, S, S, device=dev), torch.rand(B, ks, S, S, device=dev)
    ts = []
    for it in range(12):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.sepconv_forward(I, V, H, ks); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[1] * 1e-3
    by = 4.0 * (I.numel() + V.numel() + H.numel() + B * C * S * S)
    print(json.dumps({"shape": [B, C, S, S, ks], "us": round(t * 1e6, 1), "GBs": round(by / t / 1e9), "frac_hbm": round(by / t / 6553e9, 3)}))
