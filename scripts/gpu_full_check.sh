#!/bin/bash
# full GPU pass: smoke, all GPU tests, K1/K2 timings per config, whole updates, bench
TAG=${1:-full}
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_$TAG.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
timeout 600 python scripts/bench_configs.py --reps 5 > gpurun_out/configs_$TAG.log 2>&1
timeout 600 python scripts/bench_updates.py > gpurun_out/updates_$TAG.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_$TAG.log
tail -2 gpurun_out/smoke_$TAG.log; tail -12 gpurun_out/pytest_gpu_$TAG.log; cut -c1-30,60-70,100-215 gpurun_out/configs_$TAG.log; cut -c1-250 gpurun_out/updates_$TAG.log; tail -3 gpurun_out/bench_$TAG.log | cut -c1-1500
