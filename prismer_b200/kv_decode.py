"""KV-cached greedy decoding: one new token per sequence and step instead of the reference's cache-less
re-forward of the whole prefix (roberta.py:401-406, driven by prismer_caption.py:45-50 / prismer_vqa.py:51-57).

Same arithmetic as ``engine.decoder_forward`` (eval mode) restricted to the last position: the self-attention keys / values of
earlier positions come from a per-layer cache, the cross-attention K / V of the visual tokens are projected once per call
(``engine.cross_kv``).  Every product runs on the decode-time kernels of ``csrc/decode.cu`` (``prismer_skinny_linear``, ``prismer_decode_attention``, the
post-LayerNorms as 32-row ``prismer_layernorm_fwd`` launches): ~190 small launches per step instead of ~210
persistent-GEMM launches over a growing prefix.  Token ids are checked bit-exact against the reference goldens and against the
cache-less path (tests/test_kv_decode_gpu.py).
"""
from __future__ import annotations

import ctypes

import torch

from . import _C, engine, ops
from .ops import BF16, F32, _p, check


def skinny_linear(x, w16, bias=None, *, act=0, residual=None, ln=None, out_dtype=BF16):
    """y = act(x . w16^T + bias) (+ residual) for a handful of rows; with ``ln`` (a LayerNorm module) returns (y, LN(y))."""
    M, K = x.shape
    N = w16.shape[0]
    assert x.dtype == BF16 and w16.dtype == BF16 and x.stride(1) == 1 and w16.stride(1) == 1 and w16.shape[1] == K
    out = torch.empty((M, N), dtype=out_dtype, device=x.device)
    check(_C.lib().prismer_skinny_linear(
        x.data_ptr(), x.stride(0), w16.data_ptr(), w16.stride(0), _p(bias), _p(residual), residual.stride(0) if residual is not None else 0,
        out.data_ptr(), out.stride(0), int(out_dtype == F32), M, N, K, ops.ACT.get(act, act), ops._stream()), "skinny_linear")
    if ln is None:
        return out
    y, _, _ = ops.layernorm_fwd(out, ln.weight.data, ln.bias.data, ln.eps, save_stats=False)
    return out, y


def decode_attention(q, k, v, kv_bs, kv_rs, length, heads, *, k_new=None, v_new=None, k_cache=None, v_cache=None, key_mask=None, scale=None):
    """q [B, H*64] (row stride arbitrary); keys / values addressed as ptr + b*kv_bs + j*kv_rs + h*64 for j < length; optionally the new
    token's (k_new, v_new) [B, H*64] as key ``length``, appended to (k_cache, v_cache) [B, Tmax, H*64] in the same launch."""
    B, HD = q.shape
    d = HD // heads
    o = torch.empty((B, HD), dtype=BF16, device=q.device)
    check(_C.lib().prismer_decode_attention(
        q.data_ptr(), q.stride(0), k.data_ptr() if torch.is_tensor(k) else k, v.data_ptr() if torch.is_tensor(v) else v, kv_bs, kv_rs,
        length, _p(k_new), _p(v_new), k_new.stride(0) if k_new is not None else 0, _p(k_cache), _p(v_cache),
        k_cache.stride(0) if k_cache is not None else 0, k_cache.stride(1) if k_cache is not None else 0,
        _p(key_mask), key_mask.stride(0) if key_mask is not None else 0, o.data_ptr(), o.stride(0), B, heads, d,
        scale if scale is not None else d ** -0.5, ops._stream()), "decode_attention")
    return o


# Independent batch slices decoded concurrently on separate streams (parallel branches of the captured graph).  Measured on B200 at B = 32:
# 1 chain 30.7 ms per batch, 2 chains 31.8, 4 chains 34.5, 8 chains 58 -- the branches do not overlap usefully, so the default is ONE chain.
DECODE_CHAINS = 1
# The prompt is fed in one full-sequence pass that also fills the caches (KVDecoder.prefill); False = token by token through the step kernels.
PREFILL = True
_chain_streams = {}


class KVDecoder:
    """Decode state of one batch slice [lo, hi): self-attention K/V caches of every layer (+ output layer), a view of the visual K/V
    (projected once per call for the whole batch by ``engine.cross_kv``), the prompt mask."""

    def __init__(self, dec, kv, lo: int, hi: int, max_length: int, prompt_mask=None):
        cfg = dec.config
        self.dec = dec
        self.Hd, self.nh = cfg.hidden_size, cfg.num_attention_heads
        assert self.Hd // self.nh == 64, "decode kernels are written for head dim 64 (roberta-base / roberta-large)"
        dev = kv.kv_all.device
        batch = hi - lo
        self.L = len(dec.roberta.encoder.layer)
        self.kv, self.lo = kv, lo
        assert kv.S <= 320
        self.kc = torch.zeros((self.L + 1, batch, max_length, self.Hd), dtype=BF16, device=dev)
        self.vc = torch.zeros_like(self.kc)
        self.mask = torch.ones((batch, max_length), dtype=torch.int64, device=dev)
        if prompt_mask is not None:
            self.mask[:, :prompt_mask.shape[1]] = prompt_mask.to(torch.int64)
        self.B, self.Tmax = batch, max_length

    def _self_block(self, layer, li, h, t):
        at, Hd = layer.attention, self.Hd
        grp = at.self._grp
        qkv = skinny_linear(h, grp.w16, grp.b)                                                   # fused q/k/v projection of the new token
        kc, vc = self.kc[li], self.vc[li]
        o = decode_attention(qkv[:, :Hd], kc, vc, kc.stride(0), kc.stride(1), t, self.nh, k_new=qkv[:, Hd:2 * Hd], v_new=qkv[:, 2 * Hd:],
                             k_cache=kc, v_cache=vc, key_mask=self.mask)
        _, h1 = skinny_linear(o, at.output.dense.weight._c16, at.output.dense.bias.data, residual=h, ln=at.output.LayerNorm)
        return h1

    def _mlp_block(self, layer, h):
        f = skinny_linear(h, layer.intermediate.dense.weight._c16, layer.intermediate.dense.bias.data, act="gelu")
        _, h1 = skinny_linear(f, layer.output.dense.weight._c16, layer.output.dense.bias.data, residual=h, ln=layer.output.LayerNorm)
        return h1

    def prefill(self, prompt_ids: torch.Tensor):
        """All T0 prompt positions in ONE pass of the full-sequence decoder forward (``engine.decoder_forward``: causal + prompt mask,
        the tcgen05 GEMMs over B*T0 rows), whose per-layer fused q/k/v projections fill the caches; returns the fp32 logits [B, V] of the
        last prompt position.  Replaces T0 single-token passes (3 of the 19 passes of a BASE caption: ~190 launches each)."""
        Hd, T0 = self.Hd, prompt_ids.shape[1]

        def sink(li, qkv3):
            self.kc[li][:, :T0].copy_(qkv3[..., Hd:2 * Hd])
            self.vc[li][:, :T0].copy_(qkv3[..., 2 * Hd:])

        logits, _, _, _ = engine.decoder_forward(self.dec, prompt_ids, self.mask[:, :T0].contiguous(), None, None, None, save=False,
                                                 kv=self.kv, last_only=True, kv_sink=sink)
        return logits

    def step(self, ids_so_far: torch.Tensor, need_logits: bool = True):
        """Feed the token at position t = ids_so_far.shape[1] - 1 (ids_so_far: [B, t+1] int64, contiguous); returns the fp32 logits
        [B, V] of that position (None when ``need_logits`` is False: prompt tokens before the last)."""
        dec, Hd = self.dec, self.Hd
        cfg = dec.config
        t = ids_so_far.shape[1] - 1
        emb = dec.roberta.embeddings
        e, _ = ops.embed_fwd(ids_so_far, emb.word_embeddings.weight._c16, emb.position_embeddings.weight._c16,
                             emb.token_type_embeddings.weight._c16, cfg.pad_token_id)            # positions need the whole row (cumsum of non-pad)
        last = e.view(self.B, t + 1, Hd)[:, t]                                                   # [B, Hd] strided rows
        h, _, _ = ops.layernorm_fwd(last, emb.LayerNorm.weight.data, emb.LayerNorm.bias.data, emb.LayerNorm.eps, save_stats=False)
        encoder = dec.roberta.encoder
        kv = self.kv
        ld = kv.kv_all.stride(0)
        base = kv.kv_all.data_ptr() + 2 * self.lo * kv.bs * ld                                   # this slice's first batch element
        for li, (layer, cross, adp) in enumerate(encoder.layer):
            h = self._self_block(layer, li, h, t)
            q = skinny_linear(h, cross.self.query.weight._c16, cross.self.query.bias.data)
            kbase = base + 2 * (li * 2 * Hd)
            o = decode_attention(q, kbase, kbase + 2 * Hd, kv.bs * ld, kv.rs * ld, kv.S, self.nh)   # visual tokens: no mask (roberta.py:225)
            _, hc = skinny_linear(o, cross.output.dense.weight._c16, cross.output.dense.bias.data, residual=h, ln=cross.output.LayerNorm)
            a = skinny_linear(hc, adp.adaptor.down_proj.weight._c16, adp.adaptor.down_proj.bias.data, act="sqrelu")
            _, ha = skinny_linear(a, adp.adaptor.up_proj.weight._c16, adp.adaptor.up_proj.bias.data, residual=hc, ln=adp.adaptor_ln)
            h = self._mlp_block(layer, ha)
        h = self._self_block(encoder.output_layer, self.L, h, t)
        h = self._mlp_block(encoder.output_layer, h)
        if not need_logits:
            return None
        lm = dec.lm_head
        _, xl = skinny_linear(h, lm.dense.weight._c16, lm.dense.bias.data, act="gelu", ln=lm.layer_norm)
        return skinny_linear(xl, emb.word_embeddings.weight._c16, lm.bias.data, out_dtype=F32)      # tied LM head, fp32 logits [B, V]


def _chain(dec, ids, T0, kv, lo, hi, max_length, min_length, early_exit, steps, prompt_mask):
    """Greedy loop of the batch slice [lo, hi) (``ids``: that slice of the [B, max_length] id buffer) on the current stream."""
    cfg = dec.config
    eos, pad, V = cfg.eos_token_id, cfg.pad_token_id, cfg.vocab_size
    st = KVDecoder(dec, kv, lo, hi, max_length, prompt_mask)
    unfinished = torch.ones(hi - lo, dtype=torch.int64, device=ids.device)
    last = None
    if PREFILL and T0 > 1 and lo == 0 and hi == kv.B:
        last = st.prefill(ids[:, :T0].contiguous())
    else:
        for t in range(T0):
            last = st.step(ids[:, :t + 1].contiguous(), need_logits=(t == T0 - 1))
    cur = T0
    while cur < max_length:
        tok = ops.argmax(last, V, suppress_eos=cur < min_length, eos=eos)
        if steps is not None:
            steps.append(last.clone())
        tok = tok * unfinished + pad * (1 - unfinished)
        ids[:, cur] = tok
        unfinished = unfinished * (tok != eos).long()
        cur += 1
        if cur >= max_length or (early_exit and int(unfinished.max()) == 0):
            break
        last = st.step(ids[:, :cur].contiguous())
    return cur


def greedy_loop(dec, ids, T0, enc, max_length, min_length, early_exit, steps=None, prompt_mask=None, chains=None):
    """KV-cached counterpart of ``generation._greedy_loop_nocache`` (same contract): the prompt goes through one full-sequence pass that
    fills the caches (``KVDecoder.prefill``), then one token per step.  With ``early_exit=False`` there is no host synchronisation, so the loop can be captured in a
    CUDA graph.  Optionally (``chains`` / ``DECODE_CHAINS`` > 1) the batch is split into slices decoded on separate streams (parallel
    branches of the graph); measured on B200 this does not pay (see DECODE_CHAINS), so one chain is the default."""
    B = ids.shape[0]
    kv = engine.cross_kv(dec, enc)                       # visual K/V of all layers and the whole batch: once per call
    assert kv.B == B
    n = chains if chains is not None else (1 if (early_exit or steps is not None or B < 2 * DECODE_CHAINS or DECODE_CHAINS <= 1) else DECODE_CHAINS)
    if n <= 1:
        return _chain(dec, ids, T0, kv, 0, B, max_length, min_length, early_exit, steps, prompt_mask)
    main = torch.cuda.current_stream(ids.device)
    bounds = [(B * c // n, B * (c + 1) // n) for c in range(n)]
    used = []
    for c, (lo, hi) in enumerate(bounds):
        key = (str(ids.device), c)
        s = _chain_streams.get(key)
        if s is None:
            s = _chain_streams[key] = torch.cuda.Stream(device=ids.device)
        s.wait_stream(main)
        with torch.cuda.stream(s):
            _chain(dec, ids[lo:hi], T0, kv, lo, hi, max_length, min_length, False, None, None if prompt_mask is None else prompt_mask[lo:hi])
        used.append(s)
    for s in used:
        main.wait_stream(s)
    return max_length
