"""Kernel-level breakdown of ONE bench.py step (torch.profiler / CUPTI, no replay): every kernel that ran
inside the step with its launch count, summed device time and share of the step's device time.  This is the
whole-step launch list (ncu cannot serialise the cuDNN autotuning kernels of the training step reliably);
the ncu launch list of this library's own kernels is captured separately with a kernel-name filter.

    python tools/step_profile.py [--workload kth_train_b32] [--out profiles/rNN_step_kernels.csv]
"""
import argparse
import csv
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

OURS = ("tai::",)   # every kernel of libtai_b200 lives in namespace tai


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="kth_train_b32")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    step, _ = bench.make_step(args.workload, args.batch)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    rows = []
    for e in prof.key_averages():
        t = getattr(e, "device_time_total", None)
        if t is None:
            t = e.cuda_time_total
        if t > 0:
            rows.append((e.key, e.count, t))
    rows.sort(key=lambda r: -r[2])
    total = sum(r[2] for r in rows) or 1.0
    ours = sum(r[2] for r in rows if any(s in r[0] for s in OURS))
    out = open(args.out, "w", newline="") if args.out else sys.stdout
    w = csv.writer(out)
    w.writerow(["kernel", "launches", "device_us", "share_of_step_device_time", "this_library"])
    for name, n, t in rows:
        w.writerow([name[:160], n, "%.1f" % t, "%.5f" % (t / total), int(any(s in name for s in OURS))])
    w.writerow(["TOTAL", sum(r[1] for r in rows), "%.1f" % total, "1.0", ""])
    w.writerow(["THIS_LIBRARY", sum(r[1] for r in rows if any(s in r[0] for s in OURS)), "%.1f" % ours,
                "%.5f" % (ours / total), "1"])
    if args.out:
        out.close()
        print("wrote", args.out, "total device ms %.2f, this library %.2f ms" % (total / 1e3, ours / 1e3))


if __name__ == "__main__":
    main()
