#!/bin/bash
# GPU pass: smoke, parity tests, per-kernel config timings, bench. Outputs under gpurun_out/.
TAG=${1:-chk}
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_$TAG.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
timeout 600 python scripts/bench_configs.py --reps 5 > gpurun_out/configs_$TAG.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_$TAG.log
tail -3 gpurun_out/smoke_$TAG.log; tail -25 gpurun_out/pytest_gpu_$TAG.log; cat gpurun_out/configs_$TAG.log; tail -4 gpurun_out/bench_$TAG.log
