#!/bin/bash
# compute-sanitizer over the K1 / K2 kernels at small N (memcheck + racecheck + synccheck)
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 3 python scripts/bench_configs.py --scale 0.0005 --reps 1 \
      --cases c2_eval,c2_rho,c3_vb,c4_t_rho,k48d30_rho,k100d20_rho > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/sanitize_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" gpurun_out/sanitize_$tool.log | head -8
done
