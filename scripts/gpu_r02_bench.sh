#!/bin/bash
# round 2: smoke, GPU tests, the bench line (both arms), kernel timings
TAG=${1:-r02e}
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_$TAG.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
timeout 600 python scripts/bench_configs.py --reps 7 > gpurun_out/configs_$TAG.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?" >> gpurun_out/bench_$TAG.err
timeout 1500 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?" >> gpurun_out/bench_ref_$TAG.err
tail -2 gpurun_out/smoke_$TAG.log; tail -6 gpurun_out/pytest_gpu_$TAG.log | cut -c1-300
cut -c1-30,60-70,100-215 gpurun_out/configs_$TAG.log
tail -3 gpurun_out/bench_$TAG.err; grep '^{' gpurun_out/bench_$TAG.log | cut -c1-6000
tail -3 gpurun_out/bench_ref_$TAG.err; grep '^{' gpurun_out/bench_ref_$TAG.log | cut -c1-1200
