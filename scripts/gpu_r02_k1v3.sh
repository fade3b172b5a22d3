#!/bin/bash
# round 2: K1 with significance rounds -- parity tests, timings on mixture-drawn and on worst-case data, bench line
TAG=${1:-r02g}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
timeout 600 python scripts/bench_configs.py --reps 7 > gpurun_out/configs_$TAG.log 2>&1
timeout 600 python scripts/bench_configs.py --reps 7 --data normal --kernels k1 > gpurun_out/configs_normal_$TAG.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?" >> gpurun_out/bench_$TAG.err
tail -12 gpurun_out/pytest_gpu_$TAG.log | cut -c1-300
cut -c1-30,75-90,116-232 gpurun_out/configs_$TAG.log; echo NORMAL; cut -c1-30,75-90,116-232 gpurun_out/configs_normal_$TAG.log
tail -3 gpurun_out/bench_$TAG.err; grep '^{' gpurun_out/bench_$TAG.log | cut -c1-300
