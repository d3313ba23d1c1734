#!/bin/bash
# Round-2 measurement pass on ONE GPU (run under gpurun; everything lands in gpurun_out/, the files that are judged
# are copied to profiles/ afterwards).  Every command runs under its own timeout; ncu only after the same command
# has run plainly, and only on kernels of this library (selected by name).
set -u
mkdir -p gpurun_out
TAG=${1:-r02}
OURS='regex:sepconv|gates|reppad|bias_act|maxpool|unpool|upsample|l2_gdl|l2_normalize|gather_concat|gray_diff|grad_mix|slomo_|warp_|frames_to_u8'
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/${TAG}_bench_full_line.json 2> gpurun_out/${TAG}_bench_full_line.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_reference_arm.err; echo "ref rc=$?"
rm -f profiles/${TAG}_configs.jsonl
timeout 1500 python tools/measure_configs.py --tag ${TAG} 2>&1 | tail -20; cp profiles/${TAG}_configs.jsonl gpurun_out/ 2>/dev/null
timeout 900 python tools/op_sweep.py --iters 8 --out gpurun_out/${TAG}_op_sweep.csv; echo "op sweep rc=$?"
timeout 600 python tools/kernel_bench.py --cases kth160,kth,ucf,small,mid,copy,resample,stream --ref > gpurun_out/${TAG}_kernel_bench.jsonl 2> gpurun_out/${TAG}_kernel_bench.err; echo "kernel_bench rc=$?"
timeout 300 python tools/step_profile.py --out gpurun_out/${TAG}_step_kernels.csv > gpurun_out/${TAG}_step_profile.log 2>&1; echo "step_profile rc=$?"
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 2600 --csv \
    --log-file gpurun_out/${TAG}_ncu_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${TAG}_ncu_bench.log 2>&1
echo "ncu launches rc=$?"
ls -la gpurun_out | grep ${TAG}_ | head -30
