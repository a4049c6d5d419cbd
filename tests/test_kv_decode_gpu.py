"""-m gpu: KV-cached greedy decoding (prismer_b200/kv_decode.py, csrc/decode.cu) -- the decode-time kernels against fp32
PyTorch, and the cached schedule against the cache-less one (the reference's schedule, roberta.py:401-406) through the same public
``text_decoder.generate``.  The reference-golden id checks (tests/test_model_gpu.py, test_zzz_surface_golden_gpu.py, test_zzz_beam_gpu.py)
run through the cached path by default."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.helpers import GREEDY_CASES, TINY_DEC, beam_case_inputs, beam_decoder_state

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.mark.parametrize("M,N,K,act,res,ln,fp32", [
    (32, 768, 768, 0, True, True, False), (32, 2304, 768, 0, False, False, False), (32, 3072, 768, "gelu", False, False, False),
    (32, 768, 3072, 0, True, True, False), (32, 768, 768, "sqrelu", False, False, False), (32, 50265, 768, 0, False, False, True),
    (7, 1000, 256, 0, True, True, False), (96, 1024, 1024, "gelu", False, True, False), (33, 40, 64, 0, False, False, True)])
def test_skinny_linear(M, N, K, act, res, ln, fp32):
    from prismer_b200 import kv_decode
    g = torch.Generator(device="cuda").manual_seed(M + N)
    x = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    r = torch.randn(M, N, device="cuda", generator=g).to(torch.bfloat16) if res else None
    lnm = torch.nn.LayerNorm(N, eps=1e-5).cuda() if ln else None
    if ln:
        with torch.no_grad():
            lnm.weight.copy_(1 + 0.1 * torch.randn(N, device="cuda", generator=g)); lnm.bias.copy_(0.1 * torch.randn(N, device="cuda", generator=g))
    ref = x.float() @ w.float().t() + bias
    ref = {0: lambda t: t, "gelu": lambda t: F.gelu(t), "sqrelu": lambda t: F.relu(t) ** 2}[act](ref)
    if res:
        ref = ref + r.float()
    for rep in range(2):
        got = kv_decode.skinny_linear(x, w, bias, act=act, residual=r, ln=lnm, out_dtype=torch.float32 if fp32 else torch.bfloat16)
        torch.cuda.synchronize()
        out, y = got if ln else (got, None)
        assert _rel(out.float(), ref) < (2e-5 if fp32 else 4e-3), _rel(out.float(), ref)
        if ln:
            assert _rel(y.float(), F.layer_norm(out.float(), (N,), lnm.weight, lnm.bias, 1e-5)) < 4e-3


@pytest.mark.parametrize("B,H,L,new,masked", [(32, 12, 260, False, False), (32, 12, 7, True, True), (3, 4, 0, True, False), (2, 16, 320, False, False),
                                              (5, 12, 19, True, True)])
def test_decode_attention(B, H, L, new, masked):
    from prismer_b200 import kv_decode
    g = torch.Generator(device="cuda").manual_seed(B + L)
    HD, Tmax = H * 64, L + 3
    q = torch.randn(B, HD, device="cuda", generator=g).to(torch.bfloat16)
    kc = torch.randn(B, Tmax, HD, device="cuda", generator=g).to(torch.bfloat16)
    vc = torch.randn(B, Tmax, HD, device="cuda", generator=g).to(torch.bfloat16)
    kn = torch.randn(B, HD, device="cuda", generator=g).to(torch.bfloat16) if new else None
    vn = torch.randn(B, HD, device="cuda", generator=g).to(torch.bfloat16) if new else None
    mask = None
    if masked:
        mask = (torch.rand(B, Tmax, device="cuda", generator=g) > 0.3).long()
        mask[:, 0] = 1
    total = L + (1 if new else 0)
    keys = torch.cat([kc[:, :L], kn[:, None]], 1) if new else kc[:, :L]
    vals = torch.cat([vc[:, :L], vn[:, None]], 1) if new else vc[:, :L]
    s = torch.einsum("bhd,bjhd->bhj", q.float().view(B, H, 64), keys.float().view(B, total, H, 64)) / 8.0
    if masked:
        s = s.masked_fill(mask[:, None, :total] == 0, float("-inf"))
    ref = torch.einsum("bhj,bjhd->bhd", torch.softmax(s, -1), vals.float().view(B, total, H, 64)).reshape(B, HD)
    kc2, vc2 = kc.clone(), vc.clone()
    o = kv_decode.decode_attention(q, kc2, vc2, kc2.stride(0), kc2.stride(1), L, H, k_new=kn, v_new=vn, k_cache=kc2 if new else None,
                                   v_cache=vc2 if new else None, key_mask=mask)
    torch.cuda.synchronize()
    assert _rel(o.float(), ref) < 6e-3, _rel(o.float(), ref)
    if new:                                               # the new token's k / v were appended at row L, nothing else was touched
        assert torch.equal(kc2[:, L], kn) and torch.equal(vc2[:, L], vn)
        kc2[:, L], vc2[:, L] = kc[:, L], vc[:, L]
        assert torch.equal(kc2, kc) and torch.equal(vc2, vc)


@pytest.mark.parametrize("c", GREEDY_CASES, ids=[c["name"] for c in GREEDY_CASES])
def test_cached_schedule_matches_cacheless_schedule(c):
    """Same prompts (incl. right-padded ones) through both schedules: every step's logits agree to bf16 noise, and the token ids are
    identical wherever the cache-less top-1 / top-2 margin is decisive."""
    from prismer_b200 import generation, modeling
    dec = modeling.build_decoder(TINY_DEC)
    dec.load_state_dict(beam_decoder_state(dec.state_dict(), c["boost"]))
    dec.cuda().eval()
    ids, mask, enc = beam_case_inputs(c)
    enc = enc.cuda().to(torch.bfloat16)
    T0, max_len, min_len = c["T0"], c["T0"] + c["max_add"], c["T0"] + c["min_add"]
    outs = {}
    for kv in (False, True):
        generation.KV_CACHE = kv
        try:
            outs[kv] = generation.greedy(dec, ids.cuda(), enc, mask.cuda(), max_length=max_len, min_length=min_len, return_step_logits=True)
        finally:
            generation.KV_CACHE = True
    (o0, s0), (o1, s1) = outs[False], outs[True]
    n = min(len(s0), len(s1))
    same_prefix = True
    decisive = total = 0
    for t in range(n):
        if not same_prefix:
            break
        a, b = s0[t].float(), s1[t].float()
        assert _rel(b, a) < 1.5e-2, (t, _rel(b, a))
        top2 = a.topk(2, -1).values
        margin = top2[:, 0] - top2[:, 1]
        tok0, tok1 = o0[:, T0 + t], o1[:, T0 + t]
        for r in range(c["B"]):
            total += 1
            if float(margin[r]) > 5e-2:
                decisive += 1
                assert int(tok0[r]) == int(tok1[r]) or int(tok0[r]) == 1, (c["name"], r, t)
        same_prefix = bool((tok0 == tok1).all())
    print(f"{c['name']}: {n} steps compared, ids asserted at {decisive}/{total} positions, outputs identical: "
          f"{o0.shape == o1.shape and bool((o0 == o1).all())}")
    assert decisive * 2 >= total


def test_graphed_captioner_multi_chain_equals_single_chain():
    """With DECODE_CHAINS > 1 the CUDA-graphed captioner decodes the batch as concurrent slices (separate streams = parallel graph
    branches); rows are independent, so its ids must equal the single-chain ``generate`` bit for bit."""
    import random
    from prismer_b200 import generation, kv_decode, synthetic
    from tests.helpers import build_model
    experts = ["depth", "seg_coco", "obj_detection"]
    kv_decode.DECODE_CHAINS = 4                              # (default is 1: the multi-stream variant is kept, tested, not used)
    B = 2 * kv_decode.DECODE_CHAINS + 3                      # uneven slices
    m, _ = build_model(256, 2, 16, 64, experts, TINY_DEC, seed=5)
    m.eval()
    ex = synthetic.experts_to(synthetic.synth_experts(B, 64, experts, 64, 9), "cuda")
    prefix = torch.tensor([[0, 11, 12, 13]], device="cuda").repeat(B, 1)
    random.seed(1)
    try:
        cap = generation.GraphedCaptioner(m, ex, prefix, max_length=12, min_length=6)
    finally:
        kv_decode.DECODE_CHAINS = 1
    random.seed(2)
    got = cap().clone()
    random.seed(2)                                           # same instance-embedding draw as the replay above
    kv_decode.PREFILL = False                                # the slices feed the prompt token by token: compare like with like
    try:
        with torch.no_grad():
            enc = m.expert_encoder(ex).transpose(0, 1)
            want = m.text_decoder.generate(input_ids=prefix, encoder_hidden_states=enc, num_beams=1, max_length=12, min_length=6)
    finally:
        kv_decode.PREFILL = True
    torch.cuda.synchronize()
    got = generation.trim_finished(got, 4, TINY_DEC["eos_token_id"])
    L = min(got.shape[1], want.shape[1])
    assert torch.equal(got[:, :L], want[:, :L]), (got, want)
    assert cap.launches > 100
