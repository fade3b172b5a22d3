#!/bin/bash
# tuning pass: K1 mma-form timings for several (NB,NW) variants; usage: gpu_k1mma_cfg.sh TAG CASES "cfg1 cfg2 ..."
TAG=${1:-t}; CASES=${2:-c2_eval}; CFGS=${3:-"4,8 3,12 2,16"}
mkdir -p gpurun_out
for cfg in $CFGS; do
  echo "== NB,NW=$cfg"
  PMCB200_K1_FORM=mma PMCB200_K1_MMA_CFG=$cfg timeout 300 python scripts/bench_configs.py --reps 5 --kernels k1 --cases $CASES 2>&1 | grep case | cut -c1-30,105-215 | tee -a gpurun_out/cfg_k1mma_$TAG.log
done
