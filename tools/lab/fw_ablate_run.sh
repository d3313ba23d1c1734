# lab: the forward kernel with phases switched off one at a time (same scheme as vh_ablate_run.sh; timing only;
# tools/lab/fw_ablate.patch for the v3 kernel, v5_ablate.patch (-DTAI_V5_ABLATE) for the v5 kernel).
# Masks: 1 halo staging, 2 H TMA + wait, 4 V TMA + waits, 8 H slab -> registers, 16 all but one steady row per chunk.
L=video_frame_inpainting_b200/lib/libtai_b200.so
cp $L /tmp/orig.so
for k in 0 $(ls tools/lab/_build | sed -n "s/libtai_fw_\([0-9]*\).so/\1/p" | sort -n); do
  if [ $k = 0 ]; then cp /tmp/orig.so $L; else cp tools/lab/_build/libtai_fw_$k.so $L; fi
  echo "mask $k"
  timeout 300 python tools/kernel_bench.py --cases kth160,ucf24 --only fwd,fused_fwd --no-probe --iters 10 2>&1 | grep fwd | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('   ', d.get('case'), d.get('kernel'), round(d.get('ms_best',0),4), round(d.get('frac_fma_peak'),3))"
done
cp /tmp/orig.so $L
