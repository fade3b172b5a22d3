#!/bin/bash
# round 2: K1 matrix-instruction form with the transposed slice / table exp -- parity tests, timings, DIAG split
TAG=${1:-r02c}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
timeout 600 python scripts/bench_configs.py --reps 7 > gpurun_out/configs_$TAG.log 2>&1
for d in 1 2 4 6; do
  echo "DIAG=$d" >> gpurun_out/diag_$TAG.log
  PMCB200_K1_DIAG=$d timeout 120 python scripts/bench_configs.py --reps 7 --cases c2_eval 2>&1 | grep c2_eval | cut -c1-400 >> gpurun_out/diag_$TAG.log
done
tail -15 gpurun_out/pytest_gpu_$TAG.log | cut -c1-300
