#!/bin/bash
# ncu evidence for the bench command (B200_PROFILING.md recipe). Outputs under gpurun_out/; summaries are
# copied into profiles/ by scripts/summarize_profiles.py (run on the CPU box).
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 600 python scripts/bench_configs.py --reps 5 > gpurun_out/configs_$TAG.log 2>&1
# (1) every launch of the bench command with its device time
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --e2e-rows 1000000 > gpurun_out/bench_under_ncu_$TAG.log 2>&1
# (2) the dominant kernel of the bench step, full set, at the bench workload
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_mixture_eval -s 3 -c 1 \
    -o gpurun_out/k1_full_$TAG -f python bench.py --steps 2 --warmup 3 --e2e-rows 1000000 > gpurun_out/k1_full_$TAG.log 2>&1
# (3) K1 with responsibilities + K2 on the update workload (config 2 at 2e6 rows to keep the replays short)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k1_mixture_eval|k2_suffstats" -s 2 -c 2 \
    -o gpurun_out/update_full_$TAG -f python scripts/bench_configs.py --scale 0.2 --reps 1 --cases c2_rho > gpurun_out/update_full_$TAG.log 2>&1
cat gpurun_out/configs_$TAG.log
tail -3 gpurun_out/k1_full_$TAG.log gpurun_out/update_full_$TAG.log
ls -la gpurun_out
