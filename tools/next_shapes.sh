for s in 100,16384 10,1920,1080 1,3840,2160 1,7680,4320 1,64,64,64,64 1,25,160,160,48 100,16384; do
  timeout 120 python tools/prof_one.py --shape $s --steps 3 2>&1 | tail -1 | cut -c1-600
done
