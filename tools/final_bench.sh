# End-of-round measurement set (one gpurun call): GPU test suite, the bench lines filed under profiles/, per-kernel
# timings, the step profile and one ncu --set full capture of the separable-convolution kernels.
set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r02_bench_full_line.json 2> gpurun_out/r02_bench_full_line.err
timeout 600 python bench.py --workload slomo_infer_b8 > gpurun_out/r02_slomo_infer_b8.json 2> gpurun_out/r02_slomo_infer_b8.err
timeout 600 python bench.py --workload slomo_train_b4 --no-cpu-baseline > gpurun_out/r02_slomo_train_b4.json 2> gpurun_out/r02_slomo_train_b4.err
timeout 600 python bench.py --workload ucf_infer_b8 > gpurun_out/r02_ucf_infer_b8.json 2> gpurun_out/r02_ucf_infer_b8.err
timeout 300 python tools/kernel_bench.py --cases kth160,ucf,ucf24,kth5,kth --no-probe > gpurun_out/r02_kernel_bench_final.jsonl 2>&1
timeout 900 python tools/step_profile.py --out gpurun_out/r02_step_kernels.csv > gpurun_out/r02_step_profile.log 2>&1
ls -la gpurun_out | tail -12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sepconv_ -o gpurun_out/r02i_sepconv python tools/kernel_bench.py --cases kth160 --no-probe --iters 1 --warm 1 > gpurun_out/r02i_ncu.log 2>&1; echo ncu rc=$?
