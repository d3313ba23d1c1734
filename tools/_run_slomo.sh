for r in 0 1 2 3; do
  export TAI_SLOMO_V=$r
  timeout 600 python bench.py --workload slomo_infer_b8 --steps 30 --no-cpu-baseline > gpurun_out/q.json 2> gpurun_out/q.err; python - <<PY
import json
l=json.loads(open('gpurun_out/q.json').read().strip().splitlines()[-1])
print("V=$r", l['value'], l['roofline']['slomo_refine_blend_t_avg_us'], l['roofline']['slomo_interp_input_avg_us'])
PY
done
