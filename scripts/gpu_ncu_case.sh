#!/bin/bash
# usage: gpu_ncu_case.sh TAG KERNEL_REGEX CASES SCALE [SKIP] [COUNT] -- plain timings + ncu --set full capture(s)
TAG=$1; KRE=$2; CASES=$3; SCALE=${4:-0.3}; SKIP=${5:-1}; COUNT=${6:-1}
mkdir -p gpurun_out
timeout 300 python scripts/bench_configs.py --reps 5 --cases $CASES > gpurun_out/cfg_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s $SKIP -c $COUNT \
    -o gpurun_out/ncu_$TAG -f python scripts/bench_configs.py --scale $SCALE --reps 1 --cases $CASES > gpurun_out/ncu_$TAG.log 2>&1
cat gpurun_out/cfg_$TAG.log | cut -c1-300; tail -n 3 gpurun_out/ncu_$TAG.log
