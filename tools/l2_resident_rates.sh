for s in 2000,1024 100000,1024 16000,128 500000,128 8,64,64,64 100,64,64,64 1,128,128,128 10,128,128,128 6,640,480 100,640,480; do
  python tools/prof_one.py --shape $s --steps 50 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['shape'], round(d['ms']*1000,2),'us', round(d['algorithmic_gbs']), 'GB/s alg', len(d['plan']),'launches')"
done
