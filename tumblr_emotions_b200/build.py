"""In-tree build of libdeepsent.so for sm_100a (nvcc cross-compiles without a GPU).

Two libraries come out of the same objects:
  libdeepsent.so      the product: every symbol of include/deepsent.h, nothing else
  libdeepsent_dev.so  the same objects + the launch-policy overrides (runtime.cu compiled with -DDS_DEV) and the hardware probe
                      (probe.cu) declared in include/deepsent_dev.h - used only by tools/ and the tests that force a launch mode
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdeepsent.so")
LIB_DEV = os.path.join(HERE, "libdeepsent_dev.so")
SOURCES = ["runtime.cu", "comm.cu", "conv_tc.cu", "conv_bf16x3.cu", "conv_halo.cu", "split.cu", "gemm_simt.cu", "bn.cu", "pool.cu", "text.cu", "misc.cu"]
DEV_ONLY = ["probe.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers += [os.path.join(HERE, "..", "include", "deepsent.h"), os.path.join(HERE, "..", "include", "deepsent_dev.h")]
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(job):
        src, obj, extra = job
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, obj)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        return o

    jobs = [(s, s.replace(".cu", ".o"), []) for s in SOURCES]
    jobs += [("runtime.cu", "runtime_dev.o", ["-DDS_DEV"])] + [(s, s.replace(".cu", ".o"), ["-DDS_DEV"]) for s in DEV_ONLY]
    with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
        objs = list(ex.map(compile_one, jobs))
    prod = objs[:len(SOURCES)]
    dev = [o for o in prod if not o.endswith("runtime.o")] + objs[len(SOURCES):]
    for lib, members in ((LIB, prod), (LIB_DEV, dev)):
        if force or _stale(lib, members):
            cmd = [nvcc, "-shared", "-o", lib] + members + ["-lcudart", "-ldl"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
