#!/bin/bash
# ncu evidence for the four hot-path kernels (run under gpurun, 1 GPU). Usage: tools/profile_ops.sh <tag> [case]
# 1) plain run must exit 0, 2) launch list with per-launch device time, 3) --set full capture of our kernels.
set -u
TAG=${1:-r01}
CASE=${2:-allvalid}
OUT=gpurun_out
mkdir -p $OUT
CMD="python tools/op_timing.py --iters 2 --warmup 1 --case $CASE"
$CMD > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launch_$TAG.log 2>&1
$CMD > $OUT/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'_c64_kernel|_d64_kernel' -c 12 -f -o $OUT/prof_$TAG $CMD > $OUT/ncu_full_$TAG.log 2>&1
tail -2 $OUT/ncu_full_$TAG.log
ls -la $OUT | tail -8
