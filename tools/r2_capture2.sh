#!/bin/bash
# Round-2 evidence run, part 2: C2R after the staging fix, whole-plan captures (caches NOT flushed between passes).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
python tools/prof_misc.py c2r > $O/r2_plain_c2r_after.log 2>&1
$NCU -k regex:rows_c2r_kernel --launch-skip 2 -c 1 -f -o $O/r2_c2r_after python tools/prof_misc.py c2r --steps 1 > $O/r2_ncu_c2r_after.log 2>&1 || echo "ncu c2r failed"
plan() {  # name shape mode launches
  python tools/prof_one.py --shape $2 --mode $3 > $O/r2_plain_$1.log 2>&1 || { echo "plain $1 failed"; return; }
  $NCU --cache-control none -k 'regex:^(rows|cols|nd_)' --launch-skip $((3 * $4)) -c $4 -f -o $O/r2_plan_$1 python tools/prof_one.py --shape $2 --mode $3 --steps 1 > $O/r2_ncu_plan_$1.log 2>&1 || echo "ncu $1 failed"
}
plan 2d_100x640x480 100,640,480 c2c 2
plan 2d_100x640x480_r2c_half 100,640,480 half 2
plan 3d_100x64x64x64 100,64,64,64 c2c 1
plan 3d_100x64x64x64_r2c_half 100,64,64,64 half 1
plan 3d_10x128x128x128 10,128,128,128 c2c 1
plan 3d_1x256x256x256 1,256,256,256 c2c 3
plan 3d_1x512x512x512 1,512,512,512 c2c 3
cat $O/r2_plain_c2r_after.log | cut -c1-200
python tools/r2c_shapes.py > $O/r2_r2c_shapes.jsonl 2>&1; cut -c1-220 $O/r2_r2c_shapes.jsonl | tail -12
ls -la $O/r2_plan*.ncu-rep
python tools/c2r_shapes.py > $O/r2_c2r_shapes.jsonl 2>&1; cut -c1-330 $O/r2_c2r_shapes.jsonl
