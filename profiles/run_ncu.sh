#!/bin/bash
# ncu evidence for the bench workload (run under gpurun; writes into gpurun_out/).
#   1. launch list with per-launch device time (shares, not absolutes)
#   2. one --set full capture of each hot kernel (first launch after the warm-up steps)
set -x
OUT=gpurun_out
TAG=${1:-r2}
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv \
    env PC_BENCH_NO_CPU=1 PC_BENCH_CFG5_UTT=0 PC_BENCH_SHORT=1 python bench.py --steps 2 --warmup 3 > $OUT/${TAG}_ncu_bench.log 2>&1
for K in score_tc_wide_kernel accumulate_tcx_kernel fwdbwd_warp_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o $OUT/${TAG}_$K \
      env PC_BENCH_NO_CPU=1 PC_BENCH_CFG5_UTT=0 PC_BENCH_SHORT=1 python bench.py --steps 2 --warmup 3 > $OUT/${TAG}_ncu_$K.log 2>&1
done
ls -la $OUT
