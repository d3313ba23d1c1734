# lab: time the gV+gH kernel with phases switched off one at a time (timing-only: results are garbage).
# Build (no GPU needed): copy video_frame_inpainting_b200/csrc at commit 1d99e24 to a scratch directory, apply
# tools/lab/vh_ablate.patch, compile sepconv_bwd.cu with -DTAI_VH_ABLATE=<mask> and link it with the other objects of
# video_frame_inpainting_b200/build into tools/lab/_build/libtai_vh_<mask>.so (git-ignored, travels with gpurun).
# Masks: 1 gV stores (opaque predicate), 2 gH stores, 4 halo staging, 8 gO loads, 16 V TMA + waits, 32 H TMA + wait,
# 64 all but one sweep row per chunk, 128 reduce without shuffles, 256 V from registers, 512 halo taps from registers.
# Runs on the GPU box's scratch copy: the product library is overwritten there and restored at the end.
L=video_frame_inpainting_b200/lib/libtai_b200.so
cp $L /tmp/orig.so
for k in $(ls tools/lab/_build | sed -n "s/libtai_vh_\([0-9]*\).so/\1/p" | sort -n); do
  cp tools/lab/_build/libtai_vh_$k.so $L
  echo "mask $k"
  timeout 300 python tools/kernel_bench.py --cases kth160,ucf --only bwd_vh --no-probe --iters 10 2>&1 | grep bwd_vh | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('   ', d.get('case'), d.get('op'), round(d.get('ms', d.get('best_ms',0)),4), d.get('frac_fma_peak'))"
done
cp /tmp/orig.so $L
