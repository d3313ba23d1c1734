#!/bin/bash
# End-of-milestone check on one GPU box: full GPU test-suite, the default bench line, the reference arm,
# and the inference workloads.  Every command runs under its own timeout.
set -u
mkdir -p gpurun_out
TAG=${1:-chk}
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
for wl in kth_infer_b1 ucf_infer_b8 slomo_infer_b8; do
  timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_${wl}.json 2> gpurun_out/${TAG}_bench_${wl}.err; echo "$wl rc=$?"
done
python - <<'PY'
import json,glob,os
for f in sorted(glob.glob("gpurun_out/%s_bench*.json" % os.environ.get("TAG","chk"))):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    r=d.get("roofline") or {}
    print(os.path.basename(f), "value=%.2f %s ms/step=%.1f e2e=%.2f launches=%s top=%s frac=%s cpu=%s" % (
        d["value"], d["unit"], d["ms_per_step"], (d.get("e2e") or {}).get("value", float("nan")), d.get("gpu_launches"),
        r.get("kernel"), r.get("frac"), (d.get("cpu_baseline") or {}).get("value")))
PY
