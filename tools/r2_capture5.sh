#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
NCU="ncu --set full --clock-control none"
plan() {  # name shape mode launches
  python tools/prof_one.py --shape $2 --mode $3 > $O/r2_plain_$1.log 2>&1 || { echo "plain $1 failed"; return; }
  $NCU --cache-control none -k 'regex:(rows|cols|nd_|plane)' --launch-skip $((3 * $4)) -c $4 -f -o /tmp/r2_plan_$1 python tools/prof_one.py --shape $2 --mode $3 --steps 1 > $O/r2_ncu_plan_$1.log 2>&1 || echo "ncu $1 failed"
  ncu -i /tmp/r2_plan_$1.ncu-rep --page raw --csv > $O/r2_plan_$1.raw.csv 2>/dev/null
}
plan 3d_10x128x128x128_planeplan 10,128,128,128 c2c 2
plan 3d_10x128x128x128_r2c_half_planeplan 10,128,128,128 half 2
cat $O/r2_plain_3d_10x128x128x128_planeplan.log $O/r2_plain_3d_10x128x128x128_r2c_half_planeplan.log | cut -c1-200
