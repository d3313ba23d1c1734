#!/bin/bash
# End-of-round evidence pass on one GPU box (run under gpurun): GPU test-suite, bench lines (default workload,
# reference arm, inference workloads), whole-step kernel table, ncu launch list of this library's kernels in
# the bench step, ncu --set full captures of the separable-convolution and streaming kernels, kernel benches.
# Every ncu command comes after the same plain command exited 0, each under its own timeout.
set -u
mkdir -p gpurun_out
TAG=${1:-ev}
OURS='regex:sepconv|gates|reppad|replication_pad|flow_warp|slomo_|grad_mix|unpool|upsample2x|maxpool2x2|l2_gdl|bias_act|bias_grad|l2_normalize|frames_to_u8'
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
for wl in kth_infer_b1 ucf_infer_b8 slomo_infer_b8; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_${wl}.json 2> gpurun_out/${TAG}_bench_${wl}.err; echo "$wl rc=$?"
done
timeout 300 python tools/step_profile.py --out gpurun_out/${TAG}_step_kernels.csv > gpurun_out/${TAG}_step_profile.log 2>&1; echo "step_profile rc=$?"
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 4000 --csv \
    --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${TAG}_ncu_bench.log 2>&1
echo "ncu launches rc=$?"
timeout 300 python tools/kernel_bench.py --cases kth160 --iters 1 --warm 0 --no-probe > gpurun_out/${TAG}_plain_kbench.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k "$OURS" -c 6 \
    -o gpurun_out/${TAG}_sepconv -f python tools/kernel_bench.py --cases kth160 --iters 1 --warm 0 --no-probe \
    > gpurun_out/${TAG}_ncu_sepconv.log 2>&1
echo "ncu sepconv rc=$?"
timeout 120 python tools/resample_probe.py > gpurun_out/${TAG}_plain_probe.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k "$OURS" -s 28 -c 14 \
    -o gpurun_out/${TAG}_stream -f python tools/resample_probe.py > gpurun_out/${TAG}_ncu_stream.log 2>&1
echo "ncu stream rc=$?"
timeout 600 python tools/kernel_bench.py --cases kth,kth160,ucf,small,mid,resample,stream,copy --ref > gpurun_out/${TAG}_kbench.log 2>&1; echo "kbench rc=$?"
ls -la gpurun_out | grep ${TAG}_
