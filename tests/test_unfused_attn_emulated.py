"""CPU: the CALL GRAPH of the experimental batched-GEMM attention (engine._unfused_attn_fwd / _unfused_attn_bwd: operand views,
leading dimensions, (batch, head) strides, transposition flags, epilogue modes, scale) checked against torch autograd by running it
on an emulation of the three C entry points that interprets the very same arguments (pointer = storage offset of the view passed).
The kernels themselves are exercised on hardware by tests/test_experimental_gpu.py; this pins everything above the C ABI."""
import math

import pytest
import torch

from prismer_b200 import engine, ops


def _flat(t):
    return torch.tensor([], dtype=t.dtype).set_(t.untyped_storage())


def _mat(flat, off, rows, cols, ld):
    idx = off + torch.arange(rows)[:, None] * ld + torch.arange(cols)[None, :]
    return flat[idx]


def emu_gemm_batched(a, b, c, M, N, K, *, lda, ldb, ldc, trans_a=False, trans_b=False, batch_outer=1, batch_inner=1, a_bs=(0, 0),
                     b_bs=(0, 0), c_bs=(0, 0), aux=None, ldaux=0, aux_bs=(0, 0), rowvec=None, rowvec_bs=0, mode=0, alpha=1.0, force_bn=0):
    """include/prismer_sm100.h: C_i = epilogue(alpha * op(A_i) . op(B_i)^T), X_i = X + bo*bs_outer + bi*bs_inner."""
    fa, fb, fc = _flat(a), _flat(b), _flat(c)
    fx = _flat(aux) if aux is not None else None
    rv = rowvec.reshape(-1) if rowvec is not None else None
    for bo in range(batch_outer):
        for bi in range(batch_inner):
            prob = bo * batch_inner + bi
            oa = a.storage_offset() + bo * a_bs[0] + bi * a_bs[1]
            ob = b.storage_offset() + bo * b_bs[0] + bi * b_bs[1]
            oc = c.storage_offset() + bo * c_bs[0] + bi * c_bs[1]
            A = _mat(fa, oa, K, M, lda).t() if trans_a else _mat(fa, oa, M, K, lda)          # [M, K]
            B = _mat(fb, ob, K, N, ldb).t() if trans_b else _mat(fb, ob, N, K, ldb)          # [N, K]
            acc = A.float() @ B.float().t()
            if mode == 0:
                out = alpha * acc
            else:
                r = rv[prob * rowvec_bs: prob * rowvec_bs + M].float()[:, None]
                if mode == 1:
                    X = _mat(fx, aux.storage_offset() + bo * aux_bs[0] + bi * aux_bs[1], M, N, ldaux).float()
                    out = X * (acc - r) * alpha
                else:
                    out = torch.exp(alpha * acc - r)
            idx = oc + torch.arange(M)[:, None] * ldc + torch.arange(N)[None, :]
            fc[idx] = out.to(fc.dtype)


def emu_softmax_rows(s2d, Lk):
    s2d[:, :Lk] = torch.softmax(s2d[:, :Lk].float(), dim=-1).to(s2d.dtype)
    s2d[:, Lk:] = 0


def emu_attn_delta(dout_sf, o_sf, B, H, Lq, d):
    D = H * d
    prod = (dout_sf.float() * o_sf.float()).view(Lq, B, H, d).sum(-1)          # rows are (l*B + b)
    return prod.permute(1, 2, 0).reshape(B * H, Lq).contiguous()


@pytest.fixture()
def emulated(monkeypatch):
    monkeypatch.setattr(ops, "gemm_batched", emu_gemm_batched)
    monkeypatch.setattr(ops, "softmax_rows", emu_softmax_rows)
    monkeypatch.setattr(ops, "attn_delta", emu_attn_delta)


@pytest.mark.parametrize("B,H,S", [(2, 3, 20), (1, 2, 37)])
def test_unfused_attention_call_graph_matches_autograd(emulated, B, H, S):
    d, D = 64, 64 * H
    torch.manual_seed(S)
    qkv = (0.5 * torch.randn(S * B, 3 * D)).to(torch.bfloat16)
    do = torch.randn(S * B, D).to(torch.bfloat16)
    # reference: seq-first packed projections -> [B, H, S, d]
    ref = qkv.float().clone().requires_grad_(True)
    q, k, v = (ref.view(S, B, 3, H, d)[:, :, i].permute(1, 2, 0, 3) for i in range(3))
    p = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(d), dim=-1)
    o_ref = (p @ v).permute(2, 0, 1, 3).reshape(S * B, D)
    o_ref.backward(do.float())
    lse = torch.logsumexp(q @ k.transpose(-1, -2) / math.sqrt(d), dim=-1).detach().contiguous()       # [B, H, S]

    o = torch.empty(S * B, D, dtype=torch.bfloat16)
    P = engine._unfused_attn_fwd(qkv, o, B, S, H, True)
    rel = lambda a, b: float((a.float() - b.float()).norm() / b.float().norm())
    assert rel(o, o_ref.detach()) < 1e-2
    assert P.shape == (B * H, S, (S + 7) // 8 * 8) and rel(P[:, :, :S].reshape(B, H, S, S), p.detach()) < 1e-2
    for saved in (P, lse.float()):                      # backward from the saved probabilities, and from the LSE (recomputed P)
        dqkv = torch.zeros_like(qkv)
        engine._unfused_attn_bwd(do, qkv, o, saved, dqkv, B, S, H)
        for i, name in enumerate("qkv"):
            got = dqkv.view(S, B, 3, H, d)[:, :, i].float()
            want = ref.grad.view(S, B, 3, H, d)[:, :, i]
            assert rel(got, want) < 3e-2, (name, "P" if saved is P else "lse")


@pytest.mark.parametrize("B,H,Lq,Lk,d", [(2, 2, 8, 29, 96), (1, 3, 5, 12, 64)])
def test_gemm_attention_backward_general_layouts(emulated, B, H, Lq, Lk, d):
    """Resampler-shaped cross attention (resampler.py:30-31): q from its own buffer, K | V packed in another, Lq != Lk, d = 96."""
    D = H * d
    torch.manual_seed(Lq * Lk)
    qb = (0.5 * torch.randn(Lq * B, D)).to(torch.bfloat16)
    kvb = (0.5 * torch.randn(Lk * B, 2 * D)).to(torch.bfloat16)
    do = torch.randn(Lq * B, D).to(torch.bfloat16)
    qr, kvr = qb.float().clone().requires_grad_(True), kvb.float().clone().requires_grad_(True)
    q = qr.view(Lq, B, H, d).permute(1, 2, 0, 3)
    k = kvr.view(Lk, B, 2, H, d)[:, :, 0].permute(1, 2, 0, 3)
    v = kvr.view(Lk, B, 2, H, d)[:, :, 1].permute(1, 2, 0, 3)
    s_ = q @ k.transpose(-1, -2) / math.sqrt(d)
    o_ref = (torch.softmax(s_, -1) @ v).permute(2, 0, 1, 3).reshape(Lq * B, D)
    o_ref.backward(do.float())
    lse = torch.logsumexp(s_, -1).detach().contiguous()
    o = o_ref.detach().to(torch.bfloat16)
    dq, dkv = torch.zeros_like(qb), torch.zeros_like(kvb)
    engine._gemm_attn_bwd(do, o, qb, kvb[:, :D], kvb[:, D:], lse, dq, dkv[:, :D], dkv[:, D:], B, H, Lq, Lk)
    rel = lambda a, b: float((a.float() - b.float()).norm() / b.float().norm())
    assert rel(dq, qr.grad) < 3e-2 and rel(dkv, kvr.grad) < 3e-2
