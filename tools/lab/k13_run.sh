# lab: forward / backward timing at a small window (KS=13 or 25) for each library variant under tools/lab/_build
L=video_frame_inpainting_b200/lib/libtai_b200.so
cp $L /tmp/orig.so
echo baseline; timeout 200 python tools/lab/k13_run.py
for v in $(ls tools/lab/_build/*.so 2>/dev/null); do
  cp $v $L; echo $v; timeout 200 python tools/lab/k13_run.py
done
cp /tmp/orig.so $L
