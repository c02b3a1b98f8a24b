#!/bin/bash
# ncu --set full of the one-buffer row kernel (rows_ip_kernel): one launch each of three shapes; raw metric pages as CSV
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
one() {  # name shape
  python tools/prof_one.py --shape $2 > $O/r2_plain_$1.log 2>&1 || { echo "plain $1 failed"; return; }
  $NCU -k 'regex:rows_ip' --launch-skip 3 -c 1 -f -o /tmp/r2_$1 python tools/prof_one.py --shape $2 --steps 1 > $O/r2_ncu_$1.log 2>&1 || echo "ncu $1 failed"
  ncu -i /tmp/r2_$1.ncu-rep --page raw --csv > $O/r2_$1.raw.csv 2>/dev/null
}
one rows_ip_100x16384 100,16384
one rows_ip_25000x4096 25000,4096
one rows_ip_10000x16384 10000,16384
cat $O/r2_plain_rows_ip_*.log | cut -c1-220
