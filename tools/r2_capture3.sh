#!/bin/bash
# whole-plan captures, caches NOT flushed between passes; only the raw CSV page travels back (the reports are ~15 MB each)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
NCU="ncu --set full --clock-control none"
plan() {  # name shape mode launches
  python tools/prof_one.py --shape $2 --mode $3 > $O/r2_plain_$1.log 2>&1 || { echo "plain $1 failed"; return; }
  $NCU --cache-control none -k 'regex:^(rows|cols|nd_)' --launch-skip $((3 * $4)) -c $4 -f -o /tmp/r2_plan_$1 python tools/prof_one.py --shape $2 --mode $3 --steps 1 > $O/r2_ncu_plan_$1.log 2>&1 || echo "ncu $1 failed"
  ncu -i /tmp/r2_plan_$1.ncu-rep --page raw --csv > $O/r2_plan_$1.raw.csv 2>/dev/null
}
plan 2d_100x640x480 100,640,480 c2c 2
plan 2d_100x640x480_r2c_half 100,640,480 half 2
plan 3d_100x64x64x64 100,64,64,64 c2c 1
plan 3d_100x64x64x64_r2c_half 100,64,64,64 half 1
plan 3d_10x128x128x128 10,128,128,128 c2c 1
plan 3d_1x256x256x256 1,256,256,256 c2c 3
plan 3d_1x512x512x512 1,512,512,512 c2c 3
python tools/prof_misc.py c2r > $O/r2_plain_c2r_after2.log 2>&1
$NCU -k regex:rows_c2r_kernel --launch-skip 2 -c 1 -f -o /tmp/r2_c2r_after2 python tools/prof_misc.py c2r --steps 1 > $O/r2_ncu_c2r_after2.log 2>&1
ncu -i /tmp/r2_c2r_after2.ncu-rep --page raw --csv > $O/r2_c2r_after2.raw.csv 2>/dev/null
python tools/c2r_shapes.py 2>/dev/null > $O/r2_c2r_shapes_after.jsonl; cut -c1-300 $O/r2_c2r_shapes_after.jsonl
python -m pytest tests/test_gpu_half.py tests/test_gpu_jit.py tests/test_gpu_parity.py tests/test_gpu_random.py -m gpu -q -x 2>&1 | tail -5
ls -la $O/*.raw.csv
