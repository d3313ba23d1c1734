#!/bin/bash
# Round-1 profiling pass (run under gpurun).  Every ncu command is preceded by the same plain command and
# wrapped in its own timeout; kernels of this library are selected by name so that ncu never instruments
# the cuDNN autotuning kernels of the training step (ncu failed on implicit_convolve_sgemm there).
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
OURS='regex:sepconv|gates|reppad|replication_pad|flow_warp|slomo_|grad_mix|unpool'
timeout 300 python tools/step_profile.py --out gpurun_out/${TAG}_step_kernels.csv > gpurun_out/${TAG}_step_profile.log 2>&1
echo "step_profile rc=$?"
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_plain_bench.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${TAG}_ncu_bench.log 2>&1
echo "ncu launches rc=$?"
timeout 300 python tools/kernel_bench.py --cases kth --iters 1 --warm 0 --no-probe > gpurun_out/${TAG}_plain_kbench.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k "$OURS" -c 6 \
    -o gpurun_out/${TAG}_kernels -f python tools/kernel_bench.py --cases kth --iters 1 --warm 0 --no-probe \
    > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?"
tail -5 gpurun_out/${TAG}_plain_kbench.log
ls -la gpurun_out | head -30
for lab in "$@"; do :; done
if [ -x tools/lab/body_lab ]; then timeout 120 tools/lab/body_lab > gpurun_out/body_lab.log 2>&1; cat gpurun_out/body_lab.log; fi
if [ -x tools/lab/fwd4_lab ]; then timeout 120 tools/lab/fwd4_lab > gpurun_out/fwd4_lab.log 2>&1; cat gpurun_out/fwd4_lab.log; fi
if [ -x tools/lab/fwd4_timing_lab ]; then timeout 120 tools/lab/fwd4_timing_lab > gpurun_out/fwd4_timing_lab.log 2>&1; cat gpurun_out/fwd4_timing_lab.log; fi
