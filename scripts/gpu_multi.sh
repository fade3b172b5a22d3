#!/bin/bash
# 2-GPU pass: sharded PMC update (NCCL all-reduce of the statistics packet) + the bench contract at N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 scripts/pmc_sharded.py --rows 200000 --check > gpurun_out/multi_check.log 2>&1
timeout 300 $TR --master-port 29512 scripts/pmc_sharded.py --rows 10000000 > gpurun_out/multi_c5.log 2>&1
timeout 300 python scripts/pmc_sharded.py --rows 10000000 > gpurun_out/single_c5.log 2>&1
timeout 400 $TR --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1
timeout 200 $TR --master-port 29514 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_2gpu.log 2>&1
for f in multi_check multi_c5 single_c5 bench_2gpu bench_ref_2gpu; do echo "== $f"; grep -E '^\{|Error|error' gpurun_out/$f.log | cut -c1-1500; done
