#!/bin/bash
# ncu evidence for the tcgen05 Linear kernel (run under gpurun, 1 GPU). Usage: tools/profile_linear.sh <tag>
# The plain command must exit 0 first; then one --set full capture of parameter_generator (900x65536x256) and out_proj
# (900x256x32768, split-K) launches, 6-term default.
set -u
TAG=${1:-r01c}
OUT=gpurun_out
mkdir -p $OUT
for SHAPE in "900 65536 256" "900 256 32768"; do
  NAME=$(echo $SHAPE | tr ' ' 'x')
  CMD="python tools/linear_check.py --case $SHAPE 1 0 2 0"
  $CMD > $OUT/plain_linear_${NAME}_$TAG.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:'linear_bf16x3_kernel' -c 1 -f -o $OUT/prof_linear_${NAME}_$TAG $CMD > $OUT/ncu_linear_${NAME}_$TAG.log 2>&1
  tail -1 $OUT/ncu_linear_${NAME}_$TAG.log
done
ls -la $OUT | grep prof_linear
