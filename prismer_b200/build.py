"""In-tree build of ``libprismer_sm100.so`` (nvcc, sm_100a only).  ``python -m prismer_b200.build [--force]``."""
from __future__ import annotations

import glob
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
# Programmatic dependent launch (common.cuh: griddepcontrol in the GEMM / LayerNorm / colsum / attention / decode kernels, weight tiles
# prefetched towards L2 before the grid-dependency wait) is the DEFAULT build since round 2 (validated on B200: full -m gpu suite green,
# 1139 -> 1167 images/s).  PRISMER_PDL=0 builds the plain-launch variant next to it for A/B runs (select it with PRISMER_LIB).
PDL = os.environ.get("PRISMER_PDL", "1") != "0"
LIB = os.path.join(HERE, "libprismer_sm100.so" if PDL else "libprismer_sm100_nopdl.so")
OBJ_DIR = os.path.join(HERE, "build" if PDL else "build_nopdl")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-I", INCLUDE, "-I", CSRC]
if PDL:
    NVCC_FLAGS.append("-DPRISMER_PDL")


def _nvcc() -> str:
    cand = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    return cand if os.path.exists(cand) else "nvcc"


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode()); h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h")))
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_digest = _digest(hdrs)

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        stamp = obj + ".sha"
        dig = _digest([src]) + hdr_digest
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
            return obj, False
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        with open(stamp, "w") as f:
            f.write(dig)
        return obj, True

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(compile_one, srcs))
    objs = [o for o, _ in results]
    if force or any(ch for _, ch in results) or not os.path.exists(LIB):
        cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
