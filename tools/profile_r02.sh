#!/bin/bash
# Round-2 ncu evidence (run under gpurun, 1 GPU): tools/profile_r02.sh <tag> [ops|decoder|all]
#  ops     : --set full capture of the four sampling kernels at the config-1 shapes, all-valid and mixed inputs
#            (the tensors bench.py's roofline_ops leg times: racformer_b200.synthetic.make_op_inputs)
#  decoder : launch list of one decoder forward (bench.py's default workload) and a --set full capture of its MSMV / MSDA /
#            AdaptiveMixing-core / Linear launches on the decoder's own inputs
# Every profiled command is first run plain and must exit 0. Summaries: tools/ncu_raw_summary.py -> profiles/.
set -u
TAG=${1:-r02}
WHAT=${2:-all}
OUT=gpurun_out
mkdir -p $OUT
if [ "$WHAT" = ops ] || [ "$WHAT" = all ]; then
  for CASE in allvalid mixed; do
    CMD="python tools/op_timing.py --iters 1 --warmup 1 --case $CASE --ops msmv_fwd,msmv_bwd,msda_fwd,msda_bwd"
    $CMD > $OUT/${TAG}_plain_ops_$CASE.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:'_c64_|_d64_kernel' -c 8 -f -o $OUT/${TAG}_ops_$CASE $CMD > $OUT/${TAG}_ncu_ops_$CASE.log 2>&1
    tail -1 $OUT/${TAG}_ncu_ops_$CASE.log
    ncu -i $OUT/${TAG}_ops_$CASE.ncu-rep --page raw --csv > $OUT/${TAG}_ops_${CASE}_raw.csv 2>/dev/null
    python tools/ncu_raw_summary.py $OUT/${TAG}_ops_${CASE}_raw.csv > $OUT/${TAG}_ops_${CASE}_summary.json
  done
fi
if [ "$WHAT" = decoder ] || [ "$WHAT" = all ]; then
  CMD="python tools/decoder_ncu_target.py"
  $CMD > $OUT/${TAG}_plain_decoder.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/${TAG}_decoder_launches.csv $CMD > $OUT/${TAG}_ncu_decoder_launch.log 2>&1
  $CMD > $OUT/${TAG}_plain_decoder2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:'msmv_fwd|msda_fwd|adaptive_mixing_ws|linear_bf16x3_wide|row_program|sasa_attention' -c 40 -f -o $OUT/${TAG}_decoder $CMD > $OUT/${TAG}_ncu_decoder.log 2>&1
  tail -1 $OUT/${TAG}_ncu_decoder.log
  ncu -i $OUT/${TAG}_decoder.ncu-rep --page raw --csv > $OUT/${TAG}_decoder_raw.csv 2>/dev/null
  python tools/ncu_raw_summary.py $OUT/${TAG}_decoder_raw.csv > $OUT/${TAG}_decoder_summary.json
fi
# gpurun copies back at most 64 MiB: the summaries and raw CSVs are what is kept, the reports themselves only if small
for f in $(ls -S $OUT/*.ncu-rep 2>/dev/null); do
  if [ $(du -sm $OUT | cut -f1) -gt 40 ]; then rm -f $f; fi
done
ls -la $OUT | grep ${TAG}_ | tail -20
