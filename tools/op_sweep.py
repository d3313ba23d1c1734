"""BASELINE.json configs[4]: standalone SeparableConvolution op sweep (ks 13-51, 128^2-512^2 frames, batch 1-64,
C in {1,3}) forward / backward against the FP32 FMA roofline and the HBM roofline.  One CSV row per case:
the bound that applies is min(P_fma, AI x BW_hbm).

    python tools/op_sweep.py [--out profiles/rNN_op_sweep.csv] [--quick]
"""
import argparse
import csv
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_frame_inpainting_b200 import _lib, ops  # noqa: E402
from tools.kernel_bench import timeit  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    peaks = {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}
    if os.path.isfile(os.path.join(root, "MEASURED_PEAKS.json")):
        peaks.update(json.load(open(os.path.join(root, "MEASURED_PEAKS.json"))))
    dev = torch.device("cuda:0")
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    p_fma = sms * 128 * 2 * peaks["sm_max_mhz"] * 1e6
    bw = peaks["hbm_gbs"] * 1e9
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    U = lambda *s: torch.rand(*s, device=dev, generator=g) * 2 - 1
    ks_list = (13, 25, 37, 51)
    sizes = (128, 256) if args.quick else (128, 256, 512)
    batches = (1, 16) if args.quick else (1, 4, 16, 64)
    out = open(args.out, "w", newline="") if args.out else sys.stdout
    w = csv.writer(out)
    w.writerow(["ks", "size", "B", "C", "kernel", "path", "ms", "tflops", "gbs", "frac_fma_peak", "frac_hbm", "bound",
                "frac_of_bound"])
    extra = [] if args.quick else [(51, 126, 16, 1), (25, 130, 16, 3)]   # W % 4 != 0: the LDG-fed (non-TMA) kernels
    grid = [(ks, S, B, C) for ks in ks_list for S in sizes for B in batches for C in (1, 3)] + extra
    for (ks, S, B, C) in grid:
        for _once in (0,):
            for _once2 in (0,):
                for _once3 in (0,):
                    if B * ks * S * S * 4 * 4 > 40e9:   # four kernel-map sized tensors must fit comfortably
                        continue
                    I = U(B, C, S + ks - 1, S + ks - 1)
                    V, H = U(B, ks, S, S) / ks ** 0.5, U(B, ks, S, S) / ks ** 0.5
                    gO = U(B, C, S, S)
                    fl = 2.0 * B * C * S * S * ks * ks
                    px, pin = B * S * S, B * C * (S + ks - 1) ** 2
                    cases = {
                        "fwd": (lambda: ops.sepconv_forward(I, V, H, ks), fl, 4.0 * (pin + 2 * px * ks + px * C)),
                        "bwd": (lambda: ops.sepconv_backward(gO, I, V, H, ks), 3 * fl,
                                4.0 * (px * C + 2 * pin + 4 * px * ks)),
                    }
                    for name, (fn, flops, by) in cases.items():
                        med, _ = timeit(fn, iters=args.iters, warm=2, flush=flush)
                        path = _lib.last_path()
                        if name == "bwd":      # two launchers ran: ask each for its kernel family (untimed)
                            ops.sepconv_backward(gO, I, V, H, ks, (False, True, True))
                            path = _lib.last_path()
                            ops.sepconv_backward(gO, I, V, H, ks, (True, False, False))
                            path += "+" + _lib.last_path()
                        bound_t = max(flops / p_fma, by / bw)
                        bound = "fma" if flops / p_fma >= by / bw else "hbm"
                        w.writerow([ks, S, B, C, name, path, "%.4f" % (med * 1e3), "%.2f" % (flops / med / 1e12),
                                    "%.0f" % (by / med / 1e9), "%.3f" % (flops / med / p_fma), "%.3f" % (by / med / bw),
                                    bound, "%.3f" % (bound_t / med)])
                        out.flush()
                    del I, V, H, gO
    if args.out:
        out.close()


if __name__ == "__main__":
    main()
