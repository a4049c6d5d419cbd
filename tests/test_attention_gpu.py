"""-m gpu: fused attention fwd/bwd vs fp32 PyTorch (nn.MultiheadAttention core / RobertaSelfAttention math)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def _ref(q, k, v, H, causal, key_mask):
    B, Lq, HD = q.shape
    Lk, d = k.shape[1], HD // H
    qf = q.float().reshape(B, Lq, H, d).transpose(1, 2)
    kf = k.float().reshape(B, Lk, H, d).transpose(1, 2)
    vf = v.float().reshape(B, Lk, H, d).transpose(1, 2)
    s = qf @ kf.transpose(-1, -2) / math.sqrt(d)
    if causal or key_mask is not None:
        m = torch.ones(B, 1, Lq, Lk, device=q.device)
        if causal:
            m = m * torch.tril(torch.ones(Lq, Lk, device=q.device))
        if key_mask is not None:
            m = m * key_mask[:, None, None, :].float()
        s = s + (1 - m) * torch.finfo(torch.float32).min
        s = torch.max(s, torch.tensor(torch.finfo(torch.float32).min, device=q.device))
    p = torch.softmax(s, -1)
    return (p @ vf).transpose(1, 2).reshape(B, Lq, HD)


CASES = [  # B, H, Lq, Lk, d, causal, masked
    (2, 12, 260, 260, 64, False, False),   # ViT self-attention (S = 196 + 64)
    (2, 8, 64, 1240, 96, False, False),    # resampler: 64 latents over cat(latents, 6x196 expert tokens), d = 96
    (3, 12, 30, 30, 64, True, True),       # decoder causal self-attention with ragged padding
    (3, 12, 30, 260, 64, False, False),    # decoder cross-attention over the visual tokens
    (2, 8, 64, 160, 32, False, False),     # tiny-golden resampler (d = 32)
    (1, 16, 45, 1220, 64, False, False),   # LARGE VQA cross attention
    (1, 8, 64, 1600, 128, False, False),   # LARGE resampler (d = 128)
    (2, 4, 8, 8, 64, True, True),
    (2, 4, 100, 100, 64, True, False),     # causal across tile boundary
    # tcgen05 / TMEM path (csrc/attention_sm100.cu): d = 64, no mask / dropout, 64 <= Lq <= 320, Lk <= 320
    (3, 12, 320, 320, 64, False, False),   # Prismer-LARGE @224: S = 256 + 64 (16 heads there; 3 full-ish tiles)
    (2, 12, 196, 196, 64, False, False),   # PrismerZ (no latents)
    (2, 16, 256, 256, 64, False, False),   # exactly two full tiles
    (1, 4, 64, 64, 64, False, False),      # smallest shape routed to the tensor-memory kernels
    (2, 3, 130, 130, 64, False, False),    # two-row tail tile
    (2, 4, 100, 300, 64, False, False),    # Lq != Lk
    (2, 4, 272, 17, 64, False, False),     # one-row key tail, three query tiles
    (32, 12, 260, 260, 64, False, False),  # the Prismer-BASE training shape (384 CTAs on 148 SMs)
]


@pytest.mark.parametrize("B,H,Lq,Lk,d,causal,masked", CASES)
def test_attention_fwd_bwd(B, H, Lq, Lk, d, causal, masked):
    from prismer_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(Lq * 7 + Lk)
    mk = lambda L: torch.randn(B, L, H * d, device="cuda", generator=g).to(torch.bfloat16)
    # packed projections: q/k/v are strided views, as on the real path
    if Lq == Lk:
        qkv = torch.randn(B, Lq, 3 * H * d, device="cuda", generator=g).to(torch.bfloat16)
        q, k, v = qkv[..., :H * d], qkv[..., H * d:2 * H * d], qkv[..., 2 * H * d:]
    else:
        q = mk(Lq)
        kv = torch.randn(B, Lk, 2 * H * d, device="cuda", generator=g).to(torch.bfloat16)
        k, v = kv[..., :H * d], kv[..., H * d:]
    key_mask = None
    if masked:
        lens = torch.randint(1, Lk + 1, (B,), generator=torch.Generator().manual_seed(1))
        lens[0] = Lk
        key_mask = (torch.arange(Lk)[None] < lens[:, None]).long().cuda()
    o, lse = ops.attention_fwd(q, k, v, H, causal=causal, key_mask=key_mask)
    qf, kf, vf = (t.float().detach().clone().requires_grad_(True) for t in (q, k, v))
    ref = _ref(qf, kf, vf, H, causal, key_mask)
    torch.cuda.synchronize()
    assert _rel(o.float(), ref) < 4e-3, _rel(o.float(), ref)
    do = mk(Lq)
    ref.backward(do.float())
    dq, dk, dv = ops.attention_bwd(do, q, k, v, o, lse, H, causal=causal, key_mask=key_mask)
    torch.cuda.synchronize()
    assert _rel(dv.float(), vf.grad) < 8e-3, _rel(dv.float(), vf.grad)
    assert _rel(dq.float(), qf.grad) < 8e-3, _rel(dq.float(), qf.grad)
    assert _rel(dk.float(), kf.grad) < 8e-3, _rel(dk.float(), kf.grad)


def test_attention_dropout_consistency():
    """Dropout: unbiased, reproducible for a fixed (seed, stream), and the backward uses the forward's mask
    (checked through the bilinear identity <dO, O> = <dV, V>, which only holds if both used the same P_drop)."""
    from prismer_b200 import ops
    B, H, L, S, d = 2, 12, 30, 260, 64
    seed = torch.tensor([99], dtype=torch.int64, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(3)
    q = torch.randn(B, L, H * d, device="cuda", generator=g).to(torch.bfloat16)
    k = torch.randn(B, S, H * d, device="cuda", generator=g).to(torch.bfloat16)
    v = torch.ones(B, S, H * d, device="cuda", dtype=torch.bfloat16)
    o, _ = ops.attention_fwd(q, k, v, H, drop_p=0.1, seed=seed, rng_stream=5)
    o2, _ = ops.attention_fwd(q, k, v, H, drop_p=0.1, seed=seed, rng_stream=5)
    torch.cuda.synchronize()
    assert torch.equal(o, o2)
    assert abs(o.float().mean().item() - 1.0) < 2e-2
    assert o.float().std().item() > 1e-3       # dropout actually happened
    v = torch.randn(B, S, H * d, device="cuda", generator=g).to(torch.bfloat16)
    o, lse = ops.attention_fwd(q, k, v, H, drop_p=0.1, seed=seed, rng_stream=5)
    do = torch.randn(B, L, H * d, device="cuda", generator=g).to(torch.bfloat16)
    dq, dk, dv = ops.attention_bwd(do, q, k, v, o, lse, H, drop_p=0.1, seed=seed, rng_stream=5)
    torch.cuda.synchronize()
    lhs = (do.float() * o.float()).sum().item()
    rhs = (dv.float() * v.float()).sum().item()
    assert abs(lhs - rhs) / max(abs(lhs), 1.0) < 2e-2, (lhs, rhs)


@pytest.mark.parametrize("B,H,S", [(4, 12, 260), (2, 16, 320), (3, 12, 196)])
def test_tensor_memory_kernels_match_mma_sync_kernels_on_the_engine_layout(B, H, S):
    """The two implementations behind prismer_attention_fwd / _bwd (tcgen05 + TMEM, csrc/attention_sm100.cu; mma.sync,
    csrc/attention.cu) on the layout the ViT blocks use: q / k / v are column slices of the seq-first packed projection
    [S*B, 3D] (rows s*B + b), outputs are written through the same strides."""
    from prismer_b200 import engine, ops
    d, D = 64, 64 * H
    g = torch.Generator(device="cuda").manual_seed(S + B)
    qkv = torch.randn(S * B, 3 * D, device="cuda", generator=g).to(torch.bfloat16)
    do = torch.randn(S * B, D, device="cuda", generator=g).to(torch.bfloat16)
    q3 = engine._sf(qkv, S, B)
    res = []
    for legacy in (True, False):
        ops.set_attention_path(legacy)
        try:
            o = torch.zeros(S * B, D, device="cuda", dtype=torch.bfloat16)
            _, lse = ops.attention_fwd(q3[..., :D], q3[..., D:2 * D], q3[..., 2 * D:], H, out=engine._sf(o, S, B))
            dqkv = torch.zeros_like(qkv)
            d3 = engine._sf(dqkv, S, B)
            ops.attention_bwd(engine._sf(do, S, B), q3[..., :D], q3[..., D:2 * D], q3[..., 2 * D:], engine._sf(o, S, B), lse, H,
                              dq=d3[..., :D], dk=d3[..., D:2 * D], dv=d3[..., 2 * D:])
            torch.cuda.synchronize()
        finally:
            ops.set_attention_path(False)
        res.append((o, lse, dqkv))
    (o1, l1, g1), (o2, l2, g2) = res
    assert _rel(o2.float(), o1.float()) < 4e-3 and _rel(l2, l1) < 1e-5
    for c in range(3):
        assert _rel(g2[:, c * D:(c + 1) * D].float(), g1[:, c * D:(c + 1) * D].float()) < 8e-3, c


def test_shared_keys_values_equal_tiled_keys_values():
    """``kv_div`` (rank inference, prismer_caption.py:94-96 / prismer_vqa.py:95-97): k consecutive query rows read ONE set of keys / values
    -- bit-identical to attention over ``tile``-d keys / values; and the decoder with ``encoder_repeat=k`` equals the decoder on tiled
    encoder states (logits and per-sample losses)."""
    from prismer_b200 import modeling, ops
    from tests.helpers import TINY_DEC, beam_decoder_state
    B, k, H, Lq, Lk, d = 3, 5, 4, 9, 70, 64
    g = torch.Generator(device="cuda").manual_seed(11)
    q = torch.randn(B * k, Lq, H * d, device="cuda", generator=g).to(torch.bfloat16)
    kv = torch.randn(B, Lk, 2 * H * d, device="cuda", generator=g).to(torch.bfloat16)
    o1, l1 = ops.attention_fwd(q, kv[..., :H * d], kv[..., H * d:], H, kv_div=k)
    kvt = kv.repeat_interleave(k, dim=0)
    o2, l2 = ops.attention_fwd(q, kvt[..., :H * d], kvt[..., H * d:], H)
    torch.cuda.synchronize()
    assert torch.equal(o1, o2) and torch.equal(l1, l2)
    dec = modeling.build_decoder(TINY_DEC)
    dec.load_state_dict(beam_decoder_state(dec.state_dict(), 0.0))
    dec.cuda().eval()
    T, S = 7, 20
    ids = torch.randint(3, TINY_DEC["vocab_size"], (B * k, T), device="cuda", generator=g)
    mask = torch.ones_like(ids)
    labels = ids.clone(); labels[:, :3] = -100
    enc = torch.randn(B, S, TINY_DEC["vision_hidden_size"], device="cuda", generator=g).to(torch.bfloat16)
    with torch.no_grad():
        a = dec(ids, attention_mask=mask, encoder_hidden_states=enc, labels=labels, encoder_repeat=k)
        b = dec(ids, attention_mask=mask, encoder_hidden_states=enc.repeat_interleave(k, dim=0), labels=labels)
    torch.cuda.synchronize()
    assert torch.equal(a.logits, b.logits) and torch.equal(a.loss, b.loss)
