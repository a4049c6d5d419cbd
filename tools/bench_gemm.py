"""GEMM micro-benchmark over the shapes of the Prismer-BASE step (B=32): our tcgen05 kernel vs torch.matmul (cuBLAS).
CUDA-event timing, warm, L2 flushed between iterations by rotating through > 126 MB of operands."""
import json
import sys

import torch

sys.path.insert(0, ".")
from prismer_b200 import ops  # noqa: E402

SHAPES = [  # (name, M, N, K, transA, transB)
    ("vit qkv fwd", 8320, 2304, 768, 0, 0), ("vit proj fwd", 8320, 768, 768, 0, 0), ("vit fc fwd", 8320, 3072, 768, 0, 0),
    ("vit cproj fwd", 8320, 768, 3072, 0, 0), ("resampler kv fwd", 39680, 1536, 768, 0, 0), ("dec xkv grouped", 8320, 18432, 768, 0, 0),
    ("dec qkv fwd", 960, 2304, 768, 0, 0), ("dec dense fwd", 960, 768, 768, 0, 0), ("dec fc fwd", 960, 3072, 768, 0, 0),
    ("dec cproj fwd", 960, 768, 3072, 0, 0), ("lm head fwd", 960, 50265, 768, 0, 0),
    ("vit fc dgrad", 8320, 768, 3072, 0, 1), ("vit cproj dgrad", 8320, 3072, 768, 0, 1), ("lm head dgrad", 960, 768, 50265, 0, 1),
    ("adaptor wgrad", 768, 768, 8320, 1, 1), ("resampler kv wgrad", 1536, 768, 39680, 1, 1), ("xkv wgrad", 18432, 768, 8320, 1, 1),
    ("dec fc wgrad", 3072, 768, 960, 1, 1), ("dec dense wgrad", 768, 768, 960, 1, 1), ("emb wgrad", 50265, 768, 960, 1, 1),
    ("conv depth l1", 401408, 96, 16, 0, 0), ("conv depth l2", 100352, 192, 864, 0, 0), ("conv depth l3", 25088, 384, 1728, 0, 0),
    ("conv l4", 6272, 768, 3456, 0, 0), ("conv 1x1", 6272, 768, 768, 0, 0), ("conv seg l1", 25088, 96, 576, 0, 0),
    ("conv depth l2 wgrad", 192, 864, 100352, 1, 1), ("conv depth l2 dgrad", 100352, 864, 192, 0, 1),
]


def bench(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    rows = []
    for name, M, N, K, ta, tb in SHAPES:
        Kp = (K + 7) // 8 * 8
        Np = (N + 7) // 8 * 8
        a = torch.randn((K, (M + 7) // 8 * 8) if ta else (M, Kp), device="cuda").to(torch.bfloat16)
        b = torch.randn((K, Np) if tb else (N, Kp), device="cuda").to(torch.bfloat16)
        av = a[:, :M] if ta else a[:, :K]
        bv = b[:, :N] if tb else b[:, :K]
        wgrad = bool(ta and tb)      # wgrad products accumulate into fp32 (split-K eligible)
        outb = torch.zeros((M, Np), device="cuda", dtype=torch.float32 if wgrad else torch.bfloat16)[:, :N]
        res = {}
        for bn in (0, 64, 128, 256):
            try:
                res[bn] = bench(lambda: ops.gemm(av, bv, trans_a=bool(ta), trans_b=bool(tb), out=outb, force_bn=bn, accumulate=wgrad))
            except Exception as e:  # noqa
                res[bn] = float("nan")
        A = av.t() if ta else av
        Bm = bv if tb else bv.t()
        t_cublas = bench(lambda: torch.matmul(A, Bm))
        fl = 2.0 * M * N * K
        rows.append(dict(name=name, M=M, N=N, K=K, ta=ta, tb=tb, auto_ms=res[0], bn64=res[64], bn128=res[128], bn256=res[256], cublas_ms=t_cublas,
                         auto_tflops=fl / res[0] / 1e9, cublas_tflops=fl / t_cublas / 1e9))
        r = rows[-1]
        print(f"{name:22s} M{M:6d} N{N:6d} K{K:6d} t{ta}{tb} | auto {r['auto_ms']*1e3:8.1f}us {r['auto_tflops']:7.1f} TF | bn64 {res[64]*1e3:8.1f} bn128 {res[128]*1e3:8.1f} "
              f"bn256 {res[256]*1e3:8.1f} | cuBLAS {t_cublas*1e3:8.1f}us {r['cublas_tflops']:7.1f} TF", flush=True)
    json.dump(rows, open("gpurun_out/gemm_bench.json", "w"), indent=1)


if __name__ == "__main__":
    import os
    os.makedirs("gpurun_out", exist_ok=True)
    main()
