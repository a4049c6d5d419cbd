"""Round-2 experiment: ViT self-attention as batched tcgen05 GEMMs (csrc/gemm_batched_sm100.cu) against the fused mma.sync
kernels, at the BASE shape (B=32, H=12, S=260, d=64).  Checks the results first, then times each piece with CUDA events.

    timeout 120 python tools/bench_attn_unfused.py [B H S]
"""
import sys

import torch

sys.path.insert(0, ".")
from prismer_b200 import engine, ops  # noqa: E402

B, H, S = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (32, 12, 260)
d, D = 64, 64 * H
torch.manual_seed(0)
qkv = torch.randn(S * B, 3 * D, device="cuda").to(torch.bfloat16)
do = torch.randn(S * B, D, device="cuda").to(torch.bfloat16)
q3 = engine._sf(qkv, S, B)
q, k, v = q3[..., :D], q3[..., D:2 * D], q3[..., 2 * D:]
rel = lambda a, b: float((a.float() - b.float()).norm() / b.float().norm())

o_ref, lse = ops.attention_fwd(q, k, v, H)
dq, dk, dv = ops.attention_bwd(engine._sf(do, S, B), q, k, v, o_ref, lse, H)
o = torch.empty(S * B, D, device="cuda", dtype=torch.bfloat16)
P = engine._unfused_attn_fwd(qkv, o, B, S, H, True)
dqkv = torch.zeros_like(qkv)
engine._unfused_attn_bwd(do, qkv, o, P, dqkv, B, S, H)
dqkv2 = torch.zeros_like(qkv)
engine._unfused_attn_bwd(do, qkv, o, lse, dqkv2, B, S, H)
torch.cuda.synchronize()
d3, d4 = engine._sf(dqkv, S, B), engine._sf(dqkv2, S, B)
print(f"fwd O rel {rel(engine._sf(o, S, B), o_ref):.2e} | bwd(saved P) dq {rel(d3[..., :D], dq):.2e} dk {rel(d3[..., D:2 * D], dk):.2e} "
      f"dv {rel(d3[..., 2 * D:], dv):.2e} | bwd(P from LSE) dq {rel(d4[..., :D], dq):.2e} dk {rel(d4[..., D:2 * D], dk):.2e} dv {rel(d4[..., 2 * D:], dv):.2e}")


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


dog = engine._sf(do, S, B)
print(f"fused   fwd {timed(lambda: ops.attention_fwd(q, k, v, H)):8.1f} us   bwd {timed(lambda: ops.attention_bwd(dog, q, k, v, o_ref, lse, H)):8.1f} us")
print(f"unfused fwd {timed(lambda: engine._unfused_attn_fwd(qkv, o, B, S, H, True)):8.1f} us   bwd(saved P) "
      f"{timed(lambda: engine._unfused_attn_bwd(do, qkv, o, P, dqkv, B, S, H)):8.1f} us   bwd(P from LSE) "
      f"{timed(lambda: engine._unfused_attn_bwd(do, qkv, o, lse, dqkv2, B, S, H)):8.1f} us")
# pieces
Sp = P.shape[2]
ld = B * 3 * D
qs, ps, os_ = (3 * D, d), (H * S * Sp, S * Sp), (D, d)
for bn in (0, 64, 128, 256):
    t = timed(lambda: ops.gemm_batched(qkv, qkv[:, D:], P, S, S, d, lda=ld, ldb=ld, ldc=Sp, batch_outer=B, batch_inner=H, a_bs=qs, b_bs=qs,
                                       c_bs=ps, rowvec=lse, rowvec_bs=S, mode=2, alpha=d ** -0.5, force_bn=bn))
    print(f"  P = exp(scale QK^T - lse), force_bn={bn:3d}: {t:7.1f} us")
print(f"  softmax_rows: {timed(lambda: ops.softmax_rows(P.view(B * H * S, Sp), S)):7.1f} us   delta: {timed(lambda: ops.attn_delta(do, o, B, H, S, d)):7.1f} us")
print(f"  O = P V     : {timed(lambda: ops.gemm_batched(P, qkv[:, 2 * D:], o, S, d, S, lda=Sp, ldb=ld, ldc=B * D, trans_b=True, batch_outer=B, batch_inner=H, a_bs=ps, b_bs=qs, c_bs=os_)):7.1f} us")
print(f"  dV = P^T dO : {timed(lambda: ops.gemm_batched(P, do, dqkv[:, 2 * D:], S, d, S, lda=Sp, ldb=B * D, ldc=ld, trans_a=True, trans_b=True, batch_outer=B, batch_inner=H, a_bs=ps, b_bs=os_, c_bs=qs)):7.1f} us")
