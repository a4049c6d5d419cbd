"""-m gpu: the C ABI without PyTorch -- compiles examples/c_abi_gemm.cu with nvcc against include/prismer_sm100.h, links the in-tree
libprismer_sm100.so and runs it (GEMM + bias + QuickGELU and LayerNorm on cudaMalloc'd buffers vs host loops, error code for a
misaligned leading dimension)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_c_abi_example_builds_and_runs(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    from prismer_b200 import build
    build.build()
    exe = str(tmp_path / "c_abi_gemm")
    lib_dir = os.path.join(ROOT, "prismer_b200")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_abi_gemm.cu"),
           "-o", exe, "-L", lib_dir, "-lprismer_sm100", "-Xlinker", "-rpath", "-Xlinker", lib_dir]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    print(r.stdout)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr
