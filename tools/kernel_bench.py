"""Per-kernel timing of the hot-path kernels on one GPU (CUDA events on the launching stream, L2
flushed between iterations).  Prints one JSON line per case.  Development / profiling aid; the
judged numbers come from bench.py."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_frame_inpainting_b200 import ops  # noqa: E402

PEAK_FMA = 148 * 128 * 2 * 1.965e9  # nominal FP32 FMA peak, flop/s


def timeit(fn, iters=20, warm=3, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="kth,ucf,small")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--warm", type=int, default=3)
    ap.add_argument("--no-probe", action="store_true")
    ap.add_argument("--ref", action="store_true", help="also time the reference kernels (oracle/_ref)")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)  # 256 MB > 126 MB L2
    cases = {
        "kth": (32, 1, 128, 128, 51), "kth1": (1, 1, 128, 128, 51), "kth160": (160, 1, 128, 128, 51), "ucf": (8, 3, 240, 320, 51),
        "small": (16, 1, 128, 128, 13), "mid": (16, 3, 256, 256, 25),
        # launch shapes of the inference workloads (the kernel network runs once over the T*B middle frames)
        "ucf24": (24, 3, 240, 320, 51), "kth5": (5, 1, 128, 128, 51),
    }
    g = torch.Generator(device=dev).manual_seed(0)
    U = lambda *s: torch.rand(*s, device=dev, generator=g) * 2 - 1

    # FFMA ceiling
    for packed in (() if args.no_probe else (False, True)):
        grid, block, iters = 148 * 8, 256, 4096
        med, best = timeit(lambda: ops.ffma_probe(grid, block, iters, packed), iters=5)
        fl = 2.0 * 8 * iters * grid * block
        print(json.dumps({"case": "ffma_probe", "packed": packed, "tflops": fl / best / 1e12,
                          "frac_nominal": fl / best / PEAK_FMA}), flush=True)

    if "copy" in args.cases.split(","):
        # what a plain device-to-device copy reaches at the traffic volumes of the streaming kernels (the
        # MEASURED_PEAKS.json figure is a large-buffer number): total traffic = 2 x buffer size
        for mb in (16, 32, 64, 128, 256, 512, 2048):
            n = mb * 1024 * 1024 // 8   # floats per buffer: traffic (read + write) = mb MB
            src, dst = torch.randn(n, device=dev), torch.empty(n, device=dev)
            med, best = timeit(lambda: dst.copy_(src), iters=args.iters, warm=args.warm, flush=flush)
            print(json.dumps({"case": "copy", "traffic_mb": mb, "us_med": med * 1e6, "gbs": mb * 1048576 / med / 1e9,
                              "frac_hbm_6553": mb * 1048576 / med / 6553e9}), flush=True)
    if "resample" in args.cases.split(","):
        for (B, C, H, W) in [(32, 64, 64, 64), (32, 128, 32, 32), (32, 32, 64, 64), (8, 64, 120, 160)]:
            x, res, gr = U(B, C, H, W), U(B, C, 2 * H, 2 * W), U(B, C, 2 * H, 2 * W)
            out_el = B * C * 4 * H * W
            runs = {
                "upsample2x_fwd": (lambda: ops.upsample_bilinear2x_forward(x), 4.0 * (out_el + out_el / 4)),
                "upsample2x_bwd": (lambda: ops.upsample_bilinear2x_backward(gr), 4.0 * (out_el + out_el / 4)),
                "unpool_add_fwd": (lambda: ops.unpool_add_forward(x, res), 4.0 * (2 * out_el + out_el / 4)),
                "unpool_bwd": (lambda: ops.unpool_backward(gr), 4.0 * (out_el / 2)),
            }
            for k, (fn, by) in runs.items():
                med, best = timeit(fn, iters=args.iters, warm=args.warm, flush=flush)
                print(json.dumps({"case": "resample", "shape": [B, C, H, W], "kernel": k, "ms_med": med * 1e3,
                                  "gbs": by / med / 1e9, "frac_hbm_6553": by / med / 6553e9}), flush=True)
    if "stream" in args.cases.split(","):
        # ConvLSTM gates at the KTH training shape (B=32, 256 features, 16x16) and the UCF inference shape,
        # fused MSE + GDL loss at the KTH training shape ([B*T, 1, 128, 128])
        for (B, F, H, W) in [(32, 256, 16, 16), (8, 256, 30, 40)]:
            conv, state, gns = torch.randn(B, 4 * F, H, W, device=dev), torch.randn(B, 2 * F, H, W, device=dev), \
                torch.randn(B, 2 * F, H, W, device=dev)
            el = B * F * H * W
            runs = {"gates_fwd": (lambda: ops.convlstm_gates_forward(conv, state, 1.0), 28.0 * el),
                    "gates_bwd": (lambda: ops.convlstm_gates_backward(conv, state, gns, 1.0), 52.0 * el)}
            for k, (fn, by) in runs.items():
                med, best = timeit(fn, iters=args.iters, warm=args.warm, flush=flush)
                print(json.dumps({"case": "stream", "shape": [B, F, H, W], "kernel": k, "ms_med": med * 1e3,
                                  "gbs": by / med / 1e9, "frac_hbm_6553": by / med / 6553e9}), flush=True)
        for (N_, C_, H_, W_) in [(64, 64, 128, 128), (64, 128, 64, 64), (64, 256, 32, 32)]:
            y, gy, bias = torch.randn(N_, C_, H_, W_, device=dev), torch.randn(N_, C_, H_, W_, device=dev), torch.randn(C_, device=dev)
            el = y.numel()
            runs = {"bias_act_fwd": (lambda: ops.bias_act_forward_(y, bias, "relu", 0.0), 8.0 * el),
                    "bias_act_bwd": (lambda: ops.bias_act_backward(gy, y, "relu", 0.0), 12.0 * el)}
            for k, (fn, by) in runs.items():
                med, best = timeit(fn, iters=args.iters, warm=args.warm, flush=flush)
                print(json.dumps({"case": "stream", "shape": [N_, C_, H_, W_], "kernel": k, "ms_med": med * 1e3,
                                  "gbs": by / med / 1e9, "frac_hbm_6553": by / med / 6553e9}), flush=True)
        for shape in [(32, 5, 1, 128, 128), (8, 3, 3, 240, 320)]:
            x, y = U(*shape), U(*shape)
            one = torch.ones(1, device=dev)
            el = x.numel()
            runs = {"l2_gdl_fwd": (lambda: ops.l2_gdl_loss_forward(x, y), 8.0 * el),
                    "l2_gdl_bwd": (lambda: ops.l2_gdl_loss_backward(x, y, one, one), 12.0 * el)}
            for k, (fn, by) in runs.items():
                med, best = timeit(fn, iters=args.iters, warm=args.warm, flush=flush)
                print(json.dumps({"case": "stream", "shape": list(shape), "kernel": k, "ms_med": med * 1e3,
                                  "gbs": by / med / 1e9, "frac_hbm_6553": by / med / 6553e9}), flush=True)
    for name in [c for c in args.cases.split(",") if c not in ("resample", "stream", "copy")]:
        B, C, Ho, Wo, ks = cases[name]
        I = U(B, C, Ho + ks - 1, Wo + ks - 1)
        P1, P2 = U(B, C, Ho, Wo), U(B, C, Ho, Wo)
        V, H = U(B, ks, Ho, Wo) / ks ** 0.5, U(B, ks, Ho, Wo) / ks ** 0.5
        V2, H2 = U(B, ks, Ho, Wo) / ks ** 0.5, U(B, ks, Ho, Wo) / ks ** 0.5
        gO = U(B, C, Ho, Wo)
        flops = 2.0 * B * C * Ho * Wo * ks * ks
        by_fwd = 4.0 * (B * C * (Ho + ks - 1) * (Wo + ks - 1) + 2 * B * ks * Ho * Wo + B * C * Ho * Wo)
        runs = {
            "fwd": (lambda: ops.sepconv_forward(I, V, H, ks), flops, by_fwd),
            "fused_fwd": (lambda: ops.tai_fused_forward(P1, P2, V, H, V2, H2, ks), 2 * flops,
                          4.0 * (2 * B * C * Ho * Wo + 4 * B * ks * Ho * Wo + 3 * B * C * Ho * Wo)),
            "bwd_vh": (lambda: ops.sepconv_backward(gO, I, V, H, ks, (False, True, True)), 2 * flops,
                       4.0 * (B * C * Ho * Wo + B * C * (Ho + ks - 1) * (Wo + ks - 1) + 4 * B * ks * Ho * Wo)),
            "bwd_i": (lambda: ops.sepconv_backward(gO, I, V, H, ks, (True, False, False)), flops,
                      4.0 * (B * C * Ho * Wo + B * C * (Ho + ks - 1) * (Wo + ks - 1) + 2 * B * ks * Ho * Wo)),
        }
        if args.ref:
            from tests import ref_kernels
            if ref_kernels.available():
                runs["ref_fwd"] = (lambda: ref_kernels.forward(I, V, H, ks), flops, by_fwd)
                runs["ref_bwd"] = (lambda: ref_kernels.backward(gO, I, V, H, ks), 3 * flops, 0.0)
        for k, (fn, fl, by) in runs.items():
            if args.only and k not in args.only.split(","):
                continue
            med, best = timeit(fn, iters=args.iters, warm=args.warm, flush=flush)
            print(json.dumps({"case": name, "shape": [B, C, Ho, Wo, ks], "kernel": k, "ms_med": med * 1e3,
                              "ms_best": best * 1e3, "tflops": fl / med / 1e12, "frac_fma_peak": fl / med / PEAK_FMA,
                              "gbs": by / med / 1e9}), flush=True)


if __name__ == "__main__":
    main()
