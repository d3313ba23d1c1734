# lab: time the gV+gH kernel with phases switched off one at a time (libraries built from a patched copy of csrc;
# results are timing-only).  Runs on the GPU box's scratch copy: the product library is overwritten there.
L=video_frame_inpainting_b200/lib/libtai_b200.so
cp $L /tmp/orig.so
for k in 128 256 512 896; do
  cp tools/lab/_build/libtai_vh_$k.so $L
  echo "mask $k"
  timeout 300 python tools/kernel_bench.py --cases kth160,ucf --only bwd_vh --no-probe --iters 10 2>&1 | grep bwd_vh | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('   ', d.get('case'), d.get('op'), round(d.get('ms', d.get('best_ms',0)),4), d.get('frac_fma_peak'))"
done
cp /tmp/orig.so $L
