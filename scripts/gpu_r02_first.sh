#!/bin/bash
# round 2, first GPU pass: all GPU tests with the tightened tolerances (no -x: every failure and its size is wanted),
# per-config kernel timings, and ncu --set full captures of the CB=2 (C4) and CB=8 (C3) shapes of K1 / K2.
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/smi_$TAG.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -rf > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
timeout 600 python scripts/bench_configs.py --reps 7 > gpurun_out/configs_$TAG.log 2>&1
for CASE in c4_t_rho c3_vb; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k1_mma_eval|k2_suffstats" -s 1 -c 2 \
      -o gpurun_out/ncu_${TAG}_$CASE -f python scripts/bench_configs.py --scale 0.3 --reps 1 --cases $CASE > gpurun_out/ncu_${TAG}_$CASE.log 2>&1
done
for CASE in c4_t_eval c3_eval; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k1_mma_eval" -s 1 -c 1 \
      -o gpurun_out/ncu_${TAG}_$CASE -f python scripts/bench_configs.py --scale 0.3 --reps 1 --cases $CASE > gpurun_out/ncu_${TAG}_$CASE.log 2>&1
done
tail -40 gpurun_out/pytest_gpu_$TAG.log | cut -c1-400
cut -c1-30,60-70,100-215 gpurun_out/configs_$TAG.log
