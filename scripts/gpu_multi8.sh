#!/bin/bash
# N-GPU pass (N = $1): bench contract + sharded PMC iteration
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu.log 2>&1
timeout 300 $TR --master-port 29522 scripts/pmc_sharded.py --rows 10000000 > gpurun_out/multi_c5_${N}gpu.log 2>&1
timeout 200 $TR --master-port 29523 scripts/pmc_sharded.py --rows 100000 --check > gpurun_out/multi_check_${N}gpu.log 2>&1
for f in bench_${N}gpu multi_c5_${N}gpu multi_check_${N}gpu; do echo "== $f"; grep -E '^\{|Error|error' gpurun_out/$f.log | cut -c1-900; done
