#!/bin/bash
# round 2, N GPUs of one box: the NCCL test, the bench line under torchrun, the sharded scripts
TAG=${1:-r02h}; NG=${2:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi_$TAG.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -k "nccl or multi_tile or weigh or example" > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NG --steps 10 --warmup 3 > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?" >> gpurun_out/bench_$TAG.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench1_$TAG.log 2> gpurun_out/bench1_$TAG.err; echo "bench1 rc=$?" >> gpurun_out/bench1_$TAG.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29518 scripts/vb_sharded.py > gpurun_out/vb_sharded_$TAG.log 2>&1
tail -8 gpurun_out/pytest_gpu_$TAG.log | cut -c1-300
tail -4 gpurun_out/bench_$TAG.err | cut -c1-300
python - <<'PY'
import json,sys,glob
for f in sorted(glob.glob("gpurun_out/bench*_%s.log" % sys.argv[1] if len(sys.argv)>1 else "gpurun_out/bench*_r02h.log")):
    for ln in open(f):
        if ln.startswith("{"):
            d=json.loads(ln)
            print(f, d["n_gpus"], "value %.4g ms %.3f" % (d["value"], d["ms_per_step"]), "e2e", d["e2e"]["kind"], "%.4g" % d["e2e"]["value"],
                  "iter %.4g" % d["e2e_iteration"]["value"], "update", {k:v for k,v in d["update"].items() if k!="what"}, "parity", d["parity"]["max_rel_logq"])
PY
grep '^{' gpurun_out/vb_sharded_$TAG.log | cut -c1-400
