"""Round-2 experiment: cta_group::2 GEMM (csrc/gemm2_sm100.cu) against the single-CTA kernel and cuBLAS on the encoder-sized shapes
of the training step.       timeout 120 python tools/bench_gemm_2cta.py
"""
import sys

import torch

sys.path.insert(0, ".")
from prismer_b200 import ops  # noqa: E402

SHAPES = [(8320, 3072, 768), (8320, 768, 3072), (8320, 2304, 768), (8320, 768, 768), (39680, 1536, 768), (37632, 768, 768)]


def timed(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


for M, N, K in SHAPES:
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    b = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    ok = torch.equal(ops.gemm(a, b, two_cta=True), ops.gemm(a, b, two_cta=False))
    t1 = timed(lambda: ops.gemm(a, b, two_cta=False))
    t2 = timed(lambda: ops.gemm(a, b, two_cta=True))
    t2b = timed(lambda: ops.gemm(a, b, two_cta=True, force_bn=128))
    t192 = timed(lambda: ops.gemm(a, b, two_cta=False, force_bn=192)) if N % 192 == 0 else float("nan")
    tc = timed(lambda: torch.matmul(a, b.t()))
    fl = 2.0 * M * N * K / 1e6
    print(f"M{M:6d} N{N:5d} K{K:5d}: 1-CTA {t1:7.1f} us ({fl / t1:6.0f} TF/s) | 2-CTA {t2:7.1f} us ({fl / t2:6.0f}) bn128 {t2b:7.1f} us | 1-CTA bn192 {t192:7.1f} us | "
          f"cuBLAS {tc:7.1f} us ({fl / tc:6.0f}) | identical {ok}")
