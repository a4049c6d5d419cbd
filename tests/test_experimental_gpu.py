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


@pytest.mark.parametrize("M,N,K,ta,tb", [(2048, 768, 768, False, False), (8320, 3072, 768, False, False), (8320, 768, 3072, False, True),
                                           (1000, 520, 200, False, False), (4096, 768, 2048, True, True), (260, 256, 64, False, False)])
def test_two_cta_gemm_matches_single_cta(M, N, K, ta, tb):
    """cta_group::2 kernel (csrc/gemm2_sm100.cu) against the validated single-CTA kernel on the same inputs: plain, and with the
    fused bias + GELU + saved pre-activation + residual epilogue."""
    from prismer_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N)
    a = torch.randn((K, M) if ta else (M, K), device="cuda", generator=g).to(torch.bfloat16)
    b = torch.randn((K, N) if tb else (N, K), device="cuda", generator=g).to(torch.bfloat16)
    want = ops.gemm(a, b, trans_a=ta, trans_b=tb, two_cta=False)
    got = ops.gemm(a, b, trans_a=ta, trans_b=tb, two_cta=True)
    torch.cuda.synchronize()
    assert torch.equal(got, want)                       # same MMA order over K, same epilogue arithmetic -> bit-identical
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g).to(torch.bfloat16)
    z1, z2 = torch.empty_like(res), torch.empty_like(res)
    want = ops.gemm(a, b, trans_a=ta, trans_b=tb, bias=bias, act="gelu", aux_out=z1, residual=res, two_cta=False)
    got = ops.gemm(a, b, trans_a=ta, trans_b=tb, bias=bias, act="gelu", aux_out=z2, residual=res, two_cta=True)
    torch.cuda.synchronize()
    assert torch.equal(got, want) and torch.equal(z1, z2)
    f32 = ops.gemm(a, b, trans_a=ta, trans_b=tb, out_dtype=torch.float32, two_cta=True)
    assert _rel(f32, ops.gemm(a, b, trans_a=ta, trans_b=tb, out_dtype=torch.float32, two_cta=False)) < 1e-6


@pytest.mark.parametrize("rows,D,train_ln,with_res,drop", [(8320, 768, True, True, 0.0), (8320, 768, False, False, 0.0), (960, 768, True, True, 0.1),
                                                            (1030, 1024, True, False, 0.0), (77, 256, False, True, 0.0)])
def test_layernorm_bwd_v2_matches_v1(rows, D, train_ln, with_res, drop):
    """Register-lean LayerNorm backward (csrc/layernorm_v2.cu) against the validated kernel: dx / dz identical up to one bf16 ulp
    (same arithmetic), dgamma / dbeta up to fp32 summation order."""
    from prismer_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(rows + D)
    x = torch.randn(rows, D, device="cuda", generator=g).to(torch.bfloat16)
    dy = torch.randn(rows, D, device="cuda", generator=g).to(torch.bfloat16)
    gamma = torch.randn(D, device="cuda", generator=g)
    beta = torch.randn(D, device="cuda", generator=g)
    _, mean, rstd = ops.layernorm_fwd(x, gamma, beta, 1e-5, save_stats=True)
    dres = torch.randn(rows, D, device="cuda", generator=g).to(torch.bfloat16) if with_res else None
    seed = torch.tensor([1234], dtype=torch.int64, device="cuda")
    outs = []
    for v2 in (False, True):
        dg = torch.zeros(D, device="cuda") if train_ln else None
        db = torch.zeros(D, device="cuda") if train_ln else None
        dx, dz = ops.layernorm_bwd(dy, x, mean, rstd, gamma, dres=dres, dgamma=dg, dbeta=db, dz=drop > 0, drop_p=drop,
                                   seed=seed if drop > 0 else None, rng_stream=7, v2=v2)
        torch.cuda.synchronize()
        outs.append((dx, dz, dg, db))
    (dx1, dz1, dg1, db1), (dx2, dz2, dg2, db2) = outs
    ulp = lambda a, b: float(((a.float() - b.float()).abs() / b.float().abs().clamp_min(1e-2)).max())
    print(f"ln_bwd v2 vs v1 rows={rows} D={D}: dx bit-identical {bool(torch.equal(dx1, dx2))}, max rel {ulp(dx2, dx1):.2e}")
    assert ulp(dx2, dx1) <= 2 ** -7
    if dz1 is not None:
        assert ulp(dz2, dz1) <= 2 ** -7
    if train_ln:
        assert _rel(dg2, dg1) < 1e-5 and _rel(db2, db1) < 1e-5


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


@pytest.mark.parametrize("M,N,K,tb", [(8320, 768, 768, False), (8320, 768, 3072, True), (2080, 768, 2304, False), (300, 200, 136, False)])
def test_gemm_bn192_tile_matches_default(M, N, K, tb):
    """128 x 192 tile instantiation of the validated GEMM template (force_bn = 192) against the default tile choice."""
    from prismer_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + K)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    b = torch.randn((K, N) if tb else (N, K), device="cuda", generator=g).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g).to(torch.bfloat16)
    want = ops.gemm(a, b, trans_b=tb, bias=bias, act="quickgelu", residual=res)
    got = ops.gemm(a, b, trans_b=tb, bias=bias, act="quickgelu", residual=res, force_bn=192)
    torch.cuda.synchronize()
    assert torch.equal(got, want)
