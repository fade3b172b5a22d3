#!/bin/bash
# quick GPU pass: mma-form parity test + K1 timings of the mma form (cases from $2)
TAG=${1:-q}; CASES=${2:-c2_eval,c2_rho,c3_vb,c3_eval,c4_t_eval}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "matrix_instruction" > gpurun_out/pytest_k1mma_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_k1mma_$TAG.log
tail -5 gpurun_out/pytest_k1mma_$TAG.log
timeout 300 python scripts/bench_configs.py --reps 5 --kernels k1 --cases $CASES > gpurun_out/configs_k1_mma_$TAG.log 2>&1
cut -c1-30,105-250 gpurun_out/configs_k1_mma_$TAG.log
