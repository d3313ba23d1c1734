# lab: the gI kernel with phases switched off one at a time (same scheme as vh_ablate_run.sh; timing only;
# tools/lab/gi_ablate.patch holds the masks 1, 2, 4 against the quad flush).
# Masks: 1 flush body (merge + red.global), 2 flush barrier, 4 plain stores instead of atomics, 8 cross-lane reduction
# and staging (FMAs kept alive), 16 all but one steady row per chunk, 32 H slab -> registers without LDS.
L=video_frame_inpainting_b200/lib/libtai_b200.so
cp $L /tmp/orig.so
for k in 0 $(ls tools/lab/_build | sed -n "s/libtai_gi_\([0-9]*\).so/\1/p" | sort -n); do
  if [ $k = 0 ]; then cp /tmp/orig.so $L; else cp tools/lab/_build/libtai_gi_$k.so $L; fi
  echo "mask $k"
  timeout 300 python tools/kernel_bench.py --cases kth160,ucf24 --only bwd_i --no-probe --iters 10 2>&1 | grep bwd_i | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('   ', d.get('case'), d.get('kernel'), round(d.get('ms_best',0),4), round(d.get('frac_fma_peak'),3))"
done
cp /tmp/orig.so $L
