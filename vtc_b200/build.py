"""Build the CUDA library in-tree: vtc_b200/csrc/*.cu -> vtc_b200/lib/libvtc_b200.so (sm_100a).

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the
repo snapshot.  No torch headers, no libcuda link dependency (cuTensorMapEncodeTiled is resolved
at run time through cudaGetDriverEntryPoint), static cudart.
"""
from __future__ import annotations

import hashlib
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
LIB = os.path.join(LIBDIR, "libvtc_b200.so")
SOURCES = ["api.cu", "sim_tc.cu", "sim_tc_rank.cu", "sim_tc_topk.cu", "sim_tc_lse.cu", "sim_tc_store.cu",
           "exact.cu", "rank_stage.cu", "prep.cu", "reduce.cu", "cam.cu", "infonce_bwd.cu", "infonce_small.cu", "infonce_dense.cu", "cam_bwd.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _includes(path: str, seen=None):
    """Transitive closure of the quoted #include's of a source file (paths resolved like nvcc)."""
    seen = seen if seen is not None else set()
    if path in seen or not os.path.exists(path):
        return seen
    seen.add(path)
    for m in re.finditer(r'^\s*#include\s+"([^"]+)"', open(path).read(), flags=re.M):
        _includes(os.path.normpath(os.path.join(os.path.dirname(path), m.group(1))), seen)
    return seen


def _unit_hash(src: str) -> str:
    h = hashlib.sha256()
    for f in sorted(_includes(os.path.join(CSRC, src))):
        # relative names: the stamp must survive the move of the tree to the GPU box's scratch path
        h.update(os.path.relpath(f, HERE).encode())
        h.update(open(f, "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Incremental: a translation unit is recompiled only when it or a header it includes changed."""
    os.makedirs(OBJDIR, exist_ok=True)
    if not os.path.exists(NVCC):
        if os.path.exists(LIB):
            return LIB  # e.g. a box without the toolkit: use the library that travelled with the repo
        raise RuntimeError(f"nvcc not found at {NVCC}; cannot build {LIB}")

    def compile_one(src: str):
        obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        stamp = obj + ".sha256"
        want = _unit_hash(src)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == want:
            return obj, False
        cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose and r.stderr:
            print(r.stderr, file=sys.stderr)
        with open(stamp, "w") as f:
            f.write(want)
        return obj, True

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        results = list(ex.map(compile_one, SOURCES))
    objs = [o for o, _ in results]
    if not os.path.exists(LIB) or any(changed for _, changed in results):
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
