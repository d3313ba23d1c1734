"""lab: forward kernel timing at ks = 13 (HBM-bound class) for a few shapes; run once per library variant."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from video_frame_inpainting_b200 import ops
dev = torch.device("cuda:0")
flush = torch.empty(64 * 1024 * 1024, device=dev)
for (B, C, S) in [(16, 1, 128), (64, 1, 256), (64, 1, 512), (64, 3, 256), (16, 3, 512)]:
    ks = int(os.environ.get("KS", "13"))
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
    gO = torch.rand(B, C, S, S, device=dev)
    tb = []
    for it in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.sepconv_backward(gO, I, V, H, ks, (True, True, True)); e1.record(); e1.synchronize()
        tb.append(e0.elapsed_time(e1))
    tb = sorted(tb)[1] * 1e-3
    byb = 4.0 * (2 * I.numel() + 4 * V.numel() + 2 * gO.numel())
    print(json.dumps({"shape": [B, C, S, S, ks], "fwd_us": round(t * 1e6, 1), "fwd_frac_hbm": round(by / t / 6553e9, 3),
                      "bwd_us": round(tb * 1e6, 1), "bwd_frac_hbm": round(byb / tb / 6553e9, 3)}))
