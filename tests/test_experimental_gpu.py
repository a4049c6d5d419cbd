"""-m gpu, opt-in (PRISMER_EXPERIMENTAL=1): round-2 candidate kernels that are compiled and exported but not on the default
path -- the batched tcgen05 GEMM formulation of ViT self-attention.  Skipped by default so an unvalidated kernel can never take
the regular suite down."""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("PRISMER_EXPERIMENTAL") != "1", reason="opt-in experimental kernels")]


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.mark.parametrize("B,H,S", [(2, 12, 260), (1, 4, 80), (3, 12, 197)])
def test_unfused_attention_matches_fused_and_torch(B, H, S):
    from prismer_b200 import engine, ops
    d, D = 64, 64 * H
    g = torch.Generator(device="cuda").manual_seed(S)
    qkv = torch.randn(S * B, 3 * D, device="cuda", generator=g).to(torch.bfloat16)
    o = torch.empty(S * B, D, device="cuda", dtype=torch.bfloat16)
    P = engine._unfused_attn_fwd(qkv, o, B, S, H, True)
    q3 = engine._sf(qkv, S, B)
    o_ref, lse = ops.attention_fwd(q3[..., :D], q3[..., D:2 * D], q3[..., 2 * D:], H)
    torch.cuda.synchronize()
    assert _rel(engine._sf(o, S, B).float(), o_ref.float()) < 1e-2
    do = torch.randn(S * B, D, device="cuda", generator=g).to(torch.bfloat16)
    dqkv = torch.zeros_like(qkv)
    engine._unfused_attn_bwd(do, qkv, o, P, dqkv, B, S, H)
    dq, dk, dv = ops.attention_bwd(engine._sf(do, S, B), q3[..., :D], q3[..., D:2 * D], q3[..., 2 * D:], o_ref, lse, H)
    torch.cuda.synchronize()
    d3 = engine._sf(dqkv, S, B)
    assert _rel(d3[..., 2 * D:].float(), dv.float()) < 2e-2
    assert _rel(d3[..., :D].float(), dq.float()) < 2e-2
    assert _rel(d3[..., D:2 * D].float(), dk.float()) < 2e-2
    # hybrid: fused forward (LSE only) + batched-GEMM backward with P recomputed from the LSE in the score GEMM's epilogue
    dqkv2 = torch.zeros_like(qkv)
    engine._unfused_attn_bwd(do, qkv, o, lse, dqkv2, B, S, H)
    torch.cuda.synchronize()
    d4 = engine._sf(dqkv2, S, B)
    assert _rel(d4[..., 2 * D:].float(), dv.float()) < 2e-2
    assert _rel(d4[..., :D].float(), dq.float()) < 2e-2
    assert _rel(d4[..., D:2 * D].float(), dk.float()) < 2e-2


@pytest.mark.parametrize("B,H,Lq,Lk,d", [(4, 8, 64, 1240, 96), (2, 12, 30, 260, 64)])
def test_gemm_attention_backward_cross_shapes(B, H, Lq, Lk, d):
    """Resampler-shaped (Lq = 64 latents, Lk = 64 + 1176, d = 96) backward through the batched GEMMs vs the fused kernel."""
    from prismer_b200 import engine, ops
    D = H * d
    g = torch.Generator(device="cuda").manual_seed(Lk)
    qb = (0.5 * torch.randn(Lq * B, D, device="cuda", generator=g)).to(torch.bfloat16)
    kvb = (0.5 * torch.randn(Lk * B, 2 * D, device="cuda", generator=g)).to(torch.bfloat16)
    do = torch.randn(Lq * B, D, device="cuda", generator=g).to(torch.bfloat16)
    q3, kv3 = engine._sf(qb, Lq, B), engine._sf(kvb, Lk, B)
    o = torch.empty(Lq * B, D, device="cuda", dtype=torch.bfloat16)
    _, lse = ops.attention_fwd(q3, kv3[..., :D], kv3[..., D:], H, out=engine._sf(o, Lq, B))
    dq1, dkv1 = torch.empty_like(qb), torch.empty_like(kvb)
    d3 = engine._sf(dkv1, Lk, B)
    ops.attention_bwd(engine._sf(do, Lq, B), q3, kv3[..., :D], kv3[..., D:], engine._sf(o, Lq, B), lse, H, dq=engine._sf(dq1, Lq, B),
                      dk=d3[..., :D], dv=d3[..., D:])
    dq2, dkv2 = torch.zeros_like(qb), torch.zeros_like(kvb)
    engine._gemm_attn_bwd(do, o, qb, kvb[:, :D], kvb[:, D:], lse, dq2, dkv2[:, :D], dkv2[:, D:], B, H, Lq, Lk)
    torch.cuda.synchronize()
    assert _rel(dq2, dq1) < 2e-2 and _rel(dkv2, dkv1) < 2e-2
