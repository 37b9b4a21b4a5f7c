#!/bin/bash
# one --set full capture of a single kernel of the bench workload:  run_ncu_one.sh <tag> <kernel regex>
ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f -o gpurun_out/$1_$2 \
    env PC_BENCH_NO_CPU=1 PC_BENCH_CFG5_UTT=0 PC_BENCH_SHORT=1 python bench.py --steps 2 --warmup 3 > gpurun_out/$1_ncu_$2.log 2>&1
ls -la gpurun_out | tail -3
