#!/bin/bash
# round 2 diagnostics: DMMA feed microbenchmark, K1 with parts removed (PMCB200_K1_DIAG), the two tests that changed
TAG=${1:-r02b}
mkdir -p gpurun_out
timeout 300 scripts/ubench/dmma_feed > gpurun_out/dmma_feed_$TAG.log 2>&1
for d in 0 1 2 3 4 6; do
  echo "DIAG=$d" >> gpurun_out/diag_$TAG.log
  PMCB200_K1_DIAG=$d timeout 120 python scripts/bench_configs.py --reps 7 --cases c2_eval 2>&1 | grep c2_eval | cut -c1-400 >> gpurun_out/diag_$TAG.log
done
timeout 900 python -m pytest tests -m gpu -q -k "vb_fixture or multi_tile" > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
cat gpurun_out/dmma_feed_$TAG.log; cut -c1-30,100-230 gpurun_out/diag_$TAG.log; tail -5 gpurun_out/pytest_gpu_$TAG.log
