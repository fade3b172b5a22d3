#!/bin/bash
# usage: gpu_ncu_case.sh TAG KERNEL_REGEX CASES SCALE  -- plain timings + one ncu --set full capture of the kernel
TAG=$1; KRE=$2; CASES=$3; SCALE=${4:-0.3}
mkdir -p gpurun_out
timeout 600 python scripts/bench_configs.py --reps 5 --cases $CASES > gpurun_out/cfg_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s 1 -c 1 \
    -o gpurun_out/ncu_$TAG -f python scripts/bench_configs.py --scale $SCALE --reps 1 --cases $CASES > gpurun_out/ncu_$TAG.log 2>&1
cat gpurun_out/cfg_$TAG.log; tail -n 3 gpurun_out/ncu_$TAG.log
