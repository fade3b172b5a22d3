"""Build the C-ABI shared library ``pypmc_b200/csrc/libpmcb200.so`` in-tree with nvcc for sm_100a.

    python pypmc_b200/_build.py [--force]      (run as a script: importing the package would load the stale library first)

nvcc cross-compiles without a GPU.  The library travels to the GPU box with the repo snapshot; it is
git-ignored.  Objects are rebuilt only when a source or header is newer.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(CSRC, "libpmcb200.so")
SOURCES = ["pmcb200.cu", "k1_inst_0.cu", "k1_inst_1.cu", "k1_inst_2.cu", "k1_inst_3.cu", "k1_mma.cu", "k2_inst.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-I", INCLUDE,
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _newest_header() -> float:
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(INCLUDE, "pmcb200.h"))
    return max(os.path.getmtime(h) for h in hs)


def _compile(src: str, force: bool, hdr_mtime: float) -> str:
    s = os.path.join(CSRC, src)
    o = os.path.join(CSRC, src[:-3] + ".o")
    if not force and os.path.exists(o) and os.path.getmtime(o) >= max(os.path.getmtime(s), hdr_mtime):
        return o
    log = subprocess.run([_nvcc(), *NVCC_FLAGS, "-c", s, "-o", o], capture_output=True, text=True)
    with open(o[:-2] + ".ptxas.log", "w") as fh:   # registers / spills / shared memory per kernel; compile times dropped
        fh.write("".join(ln for ln in log.stderr.splitlines(True) if "Compile time" not in ln))
    if log.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (src, log.stderr[-4000:]))
    return o


def build(force: bool = False, verbose: bool = False) -> str:
    hdr = _newest_header()
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, hdr), SOURCES))
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [_nvcc(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        subprocess.check_call(cmd)
        if verbose:
            print("linked", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
