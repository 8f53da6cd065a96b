"""Build recipe for libgat.so (hand-written sm_100a CUDA + the C ABI of include/gat.h).

    python -m gpuacceleratedtracking_b200.build        # or __graft_entry__.build()

nvcc cross-compiles for sm_100a without a GPU.  The library is built IN-TREE
(gpuacceleratedtracking_b200/libgat.so) so it travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# GAT_BUILD_TAG / GAT_EXTRA_NVCC_FLAGS: A/B builds of the same sources (e.g. -DGAT_PIPE_MODE=1) into libgat_<tag>.so,
# selected at run time with GAT_LIB_PATH (scripts/dbg/ab.sh); the default build is untouched
_TAG = os.environ.get("GAT_BUILD_TAG", "")
OBJ = os.path.join(HERE, "build" + ("_" + _TAG if _TAG else ""))
LIB = os.path.join(HERE, "libgat" + ("_" + _TAG if _TAG else "") + ".so")
SOURCES = ["gat_correlate.cu", "gat_resident.cu", "gat_correlate_tc.cu", "gat_api.cu", "gat_ring.cu", "gat_mg.cu", "gat_postcorr.cu", "gat_codes.cpp"]
HEADERS = [os.path.join(CSRC, "gat_internal.h"), os.path.join(CSRC, "gat_ctx.h"), os.path.join(HERE, "..", "include", "gat.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC,-Wall", "-ccbin", "g++"] + os.environ.get("GAT_EXTRA_NVCC_FLAGS", "").split()


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src: str) -> str:
    obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
    path = os.path.join(CSRC, src)
    # gat_resident.cu includes gat_correlate.cu (the kernel body is shared)
    extra = [os.path.join(CSRC, "gat_correlate.cu")] if src == "gat_resident.cu" else []
    if _stale(obj, [path] + HEADERS + extra):
        cmd = [NVCC, *FLAGS, "-c", path, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(_compile, SOURCES))
    if force or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "g++", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
