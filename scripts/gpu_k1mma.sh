#!/bin/bash
# GPU pass for the matrix-instruction form of K1: its parity test, then per-config K1 timings for both forms.
TAG=${1:-mma}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "matrix_instruction" > gpurun_out/pytest_k1mma_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_k1mma_$TAG.log
tail -30 gpurun_out/pytest_k1mma_$TAG.log
for form in mma dfma; do
  PMCB200_K1_FORM=$form timeout 300 python scripts/bench_configs.py --reps 5 --kernels k1 > gpurun_out/configs_k1_${form}_$TAG.log 2>&1
  echo "== $form"; cut -c1-260 gpurun_out/configs_k1_${form}_$TAG.log
done
