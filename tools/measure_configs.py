"""Runs bench.py on the BASELINE.json configurations that the default line does not cover and collects the JSON
lines (one per run) into profiles/<tag>_configs.jsonl:

  config 1  KTH bi-TAI forward, batch 1: GPU (kth_infer_b1) next to the CPU port timed on the host cores
            (`cpu_baseline` of the same line and the `--impl reference` arm)
  config 3  UCF-101 bi-TAI RGB inference sweep: batch 1..32 on one GPU, and (with --gpus-list 2,4,8 on a multi-GPU
            box) batch 8 per GPU under torchrun -- clips sharded, no collective; also the reference's own 4/4/3 variant
  config 4  Super SloMo baseline
  config 5  is tools/op_sweep.py

    python tools/measure_configs.py --tag r02 [--gpus-list 1] [--quick]
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(cmd, timeout=900):
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    lines = [l for l in res.stdout.strip().splitlines() if l.startswith("{")]
    if res.returncode != 0 or not lines:
        return {"cmd": " ".join(cmd), "rc": res.returncode, "stderr_tail": res.stderr[-800:]}
    d = json.loads(lines[-1])
    d["cmd"] = " ".join(cmd[1:])
    if isinstance(d.get("roofline"), dict):
        d["roofline"].pop("kernels", None)     # keep the file small: the headline kernels stay
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="r02")
    ap.add_argument("--gpus-list", default="1")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    py = sys.executable
    out_path = os.path.join(ROOT, "profiles", "%s_configs.jsonl" % args.tag)
    gpus = [int(g) for g in args.gpus_list.split(",")]
    jobs = []
    if 1 in gpus:
        jobs.append(["bench.py", "--workload", "kth_infer_b1", "--steps", str(args.steps)])
        jobs.append(["bench.py", "--workload", "kth_infer_b1", "--impl", "reference", "--steps", "5", "--warmup", "1"])
        for b in ((1, 8) if args.quick else (1, 2, 4, 8, 16, 32)):
            jobs.append(["bench.py", "--workload", "ucf_infer_b8", "--batch", str(b), "--steps", str(args.steps)] +
                        ([] if b == 8 else ["--no-cpu-baseline"]))
        jobs.append(["bench.py", "--workload", "ucf_infer_ref443_b16", "--steps", str(args.steps), "--no-cpu-baseline"])
        jobs.append(["bench.py", "--workload", "slomo_infer_b8", "--steps", str(args.steps)])
        jobs.append(["bench.py", "--workload", "slomo_train_b4", "--steps", str(args.steps), "--no-cpu-baseline"])
    for n in gpus:
        if n > 1:
            jobs.append(["-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
                         "127.0.0.1", "--master-port", "29511", "bench.py", "--gpus", str(n), "--workload", "ucf_infer_b8",
                         "--steps", str(args.steps)])
    with open(out_path, "a") as f:
        for job in jobs:
            d = run([py] + job)
            f.write(json.dumps(d) + "\n")
            f.flush()
            r = d.get("roofline") or {}
            print("%-90s value=%s ms/step=%s e2e=%s top=%s frac=%s" % (
                d["cmd"][:90], d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), r.get("kernel"),
                r.get("frac")), flush=True)


if __name__ == "__main__":
    main()
