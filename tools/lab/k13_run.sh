L=video_frame_inpainting_b200/lib/libtai_b200.so
cp $L /tmp/orig.so
echo baseline; timeout 200 python tools/lab/k13_run.py
cp tools/lab/_build/libtai_k13.so $L
echo "min CTAs 6/4"; timeout 200 python tools/lab/k13_run.py
cp /tmp/orig.so $L
