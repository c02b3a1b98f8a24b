#!/bin/bash
# Round-2 evidence run (one B200): plain timing first, then ONE ncu --set full capture per kernel / plan, then the launch
# list of the bench command. Everything lands in gpurun_out/.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
for w in scatter slab_fused split rt c2r jit; do
  python tools/prof_misc.py $w > $O/r2_plain_$w.log 2>&1 || { echo "plain $w failed"; tail -3 $O/r2_plain_$w.log; continue; }
  case $w in
    scatter) K=cols_scatter_kernel; N=1; SKIP=2;;
    slab_fused) K=slab_fused_kernel; N=1; SKIP=2;;
    split) K=split; N=2; SKIP=4;;
    rt) K=rt_axis_kernel; N=1; SKIP=2;;
    c2r) K=rows_c2r_kernel; N=1; SKIP=2;;
    jit) K=rows_kernel; N=1; SKIP=2;;
  esac
  $NCU -k regex:$K --launch-skip $SKIP -c $N -f -o $O/r2_$w python tools/prof_misc.py $w --steps 1 > $O/r2_ncu_$w.log 2>&1 || echo "ncu $w failed"
done
# whole plans, caches NOT flushed between the passes (what one pass leaves in L2 is there for the next)
plan() {  # name shape mode launches
  python tools/prof_one.py --shape $2 --mode $3 > $O/r2_plain_$1.log 2>&1 || { echo "plain $1 failed"; return; }
  $NCU --cache-control none -k regex:b200fft --launch-skip $((3 * $4)) -c $4 -f -o $O/r2_plan_$1 python tools/prof_one.py --shape $2 --mode $3 --steps 1 > $O/r2_ncu_plan_$1.log 2>&1 || echo "ncu $1 failed"
}
plan 2d_100x640x480 100,640,480 c2c 2
plan 2d_100x640x480_r2c_half 100,640,480 half 2
plan 3d_100x64x64x64 100,64,64,64 c2c 1
plan 3d_100x64x64x64_r2c_half 100,64,64,64 half 1
plan 3d_10x128x128x128 10,128,128,128 c2c 1
plan 3d_1x256x256x256 1,256,256,256 c2c 3
plan 3d_1x512x512x512 1,512,512,512 c2c 3
cat $O/r2_plain_*.log | cut -c1-260
python bench.py --steps 20 --warmup 5 > $O/r2_bench_n1_c.json 2> $O/r2_bench_n1_c.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/r2_ncu_bench.log 2>&1; echo "launch list rc=$?"
ls -la $O/*.ncu-rep | tail -20
