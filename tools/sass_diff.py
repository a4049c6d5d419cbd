"""Per-kernel SASS comparison of the working tree against a git ref (no GPU needed): shows which device functions changed.
Used at the end of round 1 to prove that the experimental additions left every hardware-validated kernel byte-identical.

    python tools/sass_diff.py 72f6710 gemm_sm100 layernorm elementwise stems attention embed_loss
"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


def funcs(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    d, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = re.sub(r"_GLOBAL__N__[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]+", "ANON", m.group(1))
            d[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            d[cur].append(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line))
    return d


def main():
    ref, names = sys.argv[1], sys.argv[2:]
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(f"{tmp}/csrc"); os.makedirs(f"{tmp}/include")
        for path, dst in [("include/prismer_sm100.h", f"{tmp}/include/prismer_sm100.h")] + \
                         [(f"prismer_b200/csrc/{f}", f"{tmp}/csrc/{f}") for f in ("common.cuh", "sm100_ptx.cuh")] + \
                         [(f"prismer_b200/csrc/{n}.cu", f"{tmp}/csrc/{n}.cu") for n in names]:
            with open(dst, "w") as f:
                f.write(subprocess.run(["git", "show", f"{ref}:{path}"], capture_output=True, text=True, cwd=ROOT, check=True).stdout)
        for n in names:
            subprocess.run(["nvcc"] + FLAGS + ["-I", f"{tmp}/include", "-I", f"{tmp}/csrc", "-c", f"{tmp}/csrc/{n}.cu", "-o", f"{tmp}/old_{n}.o"], check=True)
            subprocess.run(["nvcc"] + FLAGS + ["-I", f"{ROOT}/include", "-I", f"{ROOT}/prismer_b200/csrc", "-c", f"{ROOT}/prismer_b200/csrc/{n}.cu",
                            "-o", f"{tmp}/new_{n}.o"], check=True)
            old, new = funcs(f"{tmp}/old_{n}.o"), funcs(f"{tmp}/new_{n}.o")
            changed = [k for k in old if old[k] != new.get(k)]
            added = [k for k in new if k not in old]
            print(f"{n}: {len(old)} kernels at {ref}, {len(changed)} changed, {len(added)} added" + (f"  CHANGED: {changed}" if changed else ""))


if __name__ == "__main__":
    main()
