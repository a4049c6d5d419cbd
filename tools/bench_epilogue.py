"""Epilogue-variant timing of the GEMM on the ViT MLP shapes (what the fused epilogues cost on top of the plain GEMM)."""
import sys
import torch
sys.path.insert(0, ".")
from prismer_b200 import ops
from tools.bench_gemm import bench

for (M, N, K) in [(8320, 3072, 768), (8320, 768, 768), (8320, 768, 3072), (960, 768, 768)]:
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    b = torch.randn(N, K, device="cuda").to(torch.bfloat16) * 0.05
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda").to(torch.bfloat16)
    aux = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    z = torch.randn(M, N, device="cuda").to(torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    seed = torch.tensor([1], dtype=torch.int64, device="cuda")
    variants = {
        "plain": dict(),
        "bias": dict(bias=bias),
        "bias+quickgelu": dict(bias=bias, act="quickgelu"),
        "bias+quickgelu+aux": dict(bias=bias, act="quickgelu", aux_out=aux),
        "bias+gelu+aux": dict(bias=bias, act="gelu", aux_out=aux),
        "bias+sqrelu+aux": dict(bias=bias, act="sqrelu", aux_out=aux),
        "bias+residual": dict(bias=bias, residual=res),
        "actgrad(quickgelu)": dict(act_grad="quickgelu", aux_in=z),
        "bias+drop+residual": dict(bias=bias, residual=res, drop_p=0.1, seed=seed, rng_stream=3),
    }
    fl = 2.0 * M * N * K
    line = f"M{M} N{N} K{K}: "
    for name, kw in variants.items():
        t = bench(lambda: ops.gemm(a, b, out=out, **kw))
        line += f"{name} {t*1e3:.1f}us | "
    print(line, flush=True)
