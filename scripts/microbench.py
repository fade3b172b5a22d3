"""FP64 throughput probes on the current GPU (run under gpurun): DFMA only, DFMA + broadcast LDS.128 mixes, DMMA."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from pypmc_b200 import _lib  # noqa: E402

ctx = _lib.Context.get(0)
names = {0: "dfma_only", 1: "dfma_lds128_per4", 2: "dfma_lds128_per2", 3: "dmma_m8n8k4"}
out = {}
for which, name in names.items():
    for iters in (2000, 20000):
        g, ms = ctx.fp64_peak(which, iters)
        out["%s_iters%d" % (name, iters)] = {"gflops": g, "ms": ms}
print(json.dumps(out, indent=1))
