#!/bin/bash
# ncu evidence for the bench command (B200_PROFILING.md recipe). Outputs under gpurun_out/; summaries are
# written into profiles/ by scripts/summarize_profiles.py (run on the CPU box, which has ncu but no GPU).
# usage: gpu_profile.sh TAG [timings|bench|shapes|shapes_eval]   (gpurun returns at most 64 MiB: one part per call)
TAG=${1:-r02}; PART=${2:-timings}
mkdir -p gpurun_out
if [ "$PART" = "timings" ]; then
timeout 300 python scripts/bench_configs.py --reps 7 > gpurun_out/configs_$TAG.log 2>&1
timeout 300 python scripts/bench_configs.py --reps 7 --data normal --kernels k1 > gpurun_out/configs_normal_$TAG.log 2>&1
timeout 600 python scripts/bench_updates.py > gpurun_out/updates_$TAG.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_$TAG.log 2>&1
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2>&1
cat gpurun_out/configs_$TAG.log | cut -c1-30,75-90,116-232
cat gpurun_out/updates_$TAG.log | cut -c1-250
grep '^{' gpurun_out/bench_$TAG.log | cut -c1-400
grep '^{' gpurun_out/bench_ref_$TAG.log | cut -c1-400
fi
if [ "$PART" = "bench" ]; then
# (1) every launch of the bench command with its device time
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 4 --warmup 3 --e2e-rows 1000000 --parity-rows 100000 > gpurun_out/bench_under_ncu_$TAG.log 2>&1
# (2) the dominant kernel of the bench step, full set, at the bench workload (N = 1e7)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_mma_eval -s 3 -c 1 \
    -o gpurun_out/k1_full_$TAG -f python bench.py --steps 2 --warmup 3 --e2e-rows 1000000 --parity-rows 100000 > gpurun_out/k1_full_$TAG.log 2>&1
# (3) K1 with responsibilities + K2 on the update workload (config 2, N = 1e7)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k1_mma_eval|k2_suffstats" -s 1 -c 2 \
    -o gpurun_out/update_full_$TAG -f python scripts/bench_configs.py --reps 1 --cases c2_rho > gpurun_out/update_full_$TAG.log 2>&1
tail -n 2 gpurun_out/k1_full_$TAG.log gpurun_out/update_full_$TAG.log
fi
if [ "$PART" = "shapes" ]; then
# (4) the CB = 2 (C4) and CB = 8 (C3) shapes, fused second pass + K2
for CASE in c4_t_rho c3_vb; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k1_mma_eval|k2_suffstats" -s 1 -c 2 \
      -o gpurun_out/ncu_${TAG}_$CASE -f python scripts/bench_configs.py --scale 0.3 --reps 1 --cases $CASE > gpurun_out/ncu_${TAG}_$CASE.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
fi
if [ "$PART" = "shapes_eval" ]; then
for CASE in c4_t_eval c3_eval; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k1_mma_eval" -s 1 -c 1 \
      -o gpurun_out/ncu_${TAG}_$CASE -f python scripts/bench_configs.py --scale 0.3 --reps 1 --cases $CASE > gpurun_out/ncu_${TAG}_$CASE.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
fi
