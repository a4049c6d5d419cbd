"""TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the Prismer hot path.

This file is the *checker*, never the product: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The product
path (``prismer_b200``) never routes through it and fails loudly without its CUDA library.

It restates, in plain functional fp32/fp64 PyTorch on CPU, the arithmetic of the reference's
forward / loss / greedy-generate path from a *reference-layout* ``state_dict``
(SURVEY.md section 8b), each function citing the reference ``file:line`` it follows.

Pinning: the reference has no golden vectors for this path, so the
oracle is pinned against outputs of the reference's own modules run in the build container:
``oracle/gen_golden.py`` imports ``/root/reference`` (through ``oracle/reference_shim.py``),
and commits small fixtures under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks
this file against them (and against the live reference when it is mounted).
"""
from __future__ import annotations

import math
import random
from typing import Dict, Optional

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# model/modules/utils.py
# --------------------------------------------------------------------------------------


def layer_norm(x, w, b, eps=1e-5):
    """utils.py:14-19 -- fp32 LayerNorm, result cast back to the input dtype."""
    return F.layer_norm(x.float(), (x.shape[-1],), w.float(), b.float(), eps).to(x.dtype)


def quick_gelu(x):
    """utils.py:23-25"""
    return x * torch.sigmoid(1.702 * x)


def squared_relu(x):
    """utils.py:28-30"""
    return torch.square(torch.relu(x))


def gelu_erf(x):
    """transformers ACT2FN['gelu'] / ``gelu`` (roberta.py:17,164,423): exact erf GELU."""
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def interpolate_pos_embed(pos, target_len):
    """utils.py:34-44 -- bicubic (align_corners=False) resize of the sqrt(P) x sqrt(P) grid."""
    o = int(pos.shape[0] ** 0.5)
    n = int(target_len ** 0.5)
    if o == n:
        return pos
    p = pos.reshape(1, o, o, -1).permute(0, 3, 1, 2)
    p = F.interpolate(p, size=(n, n), mode="bicubic", align_corners=False)
    return p.permute(0, 2, 3, 1).flatten(0, 2)


def linear(x, w, b=None):
    return F.linear(x, w, b)


def adaptor(x, sd, pre, norm_late):
    """utils.py:48-65 -- Adaptor; ``pre`` is the key prefix of the module holding
    ``adaptor.down_proj / adaptor.up_proj / adaptor_ln``."""
    def f(h):
        h = linear(h, sd[pre + "adaptor.down_proj.weight"], sd[pre + "adaptor.down_proj.bias"])
        h = squared_relu(h)
        return linear(h, sd[pre + "adaptor.up_proj.weight"], sd[pre + "adaptor.up_proj.bias"])
    lw, lb = sd[pre + "adaptor_ln.weight"], sd[pre + "adaptor_ln.bias"]
    if norm_late:
        return layer_norm(f(x) + x, lw, lb)
    return f(layer_norm(x, lw, lb)) + x


def mha(q_in, kv_in, sd, pre, heads):
    """``nn.MultiheadAttention`` packed in-proj math as used at vit.py:41,52-53 and
    resampler.py:18,31 (seq-first [L,B,D]; no mask, no dropout)."""
    w, b = sd[pre + "in_proj_weight"], sd[pre + "in_proj_bias"]
    D = q_in.shape[-1]
    q = linear(q_in, w[:D], b[:D])
    k = linear(kv_in, w[D:2 * D], b[D:2 * D])
    v = linear(kv_in, w[2 * D:], b[2 * D:])
    Lq, B, _ = q.shape
    Lk = k.shape[0]
    d = D // heads
    q = q.reshape(Lq, B, heads, d).permute(1, 2, 0, 3)
    k = k.reshape(Lk, B, heads, d).permute(1, 2, 0, 3)
    v = v.reshape(Lk, B, heads, d).permute(1, 2, 0, 3)
    s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(d)
    p = torch.softmax(s, dim=-1)
    o = torch.matmul(p, v).permute(2, 0, 1, 3).reshape(Lq, B, D)
    return linear(o, sd[pre + "out_proj.weight"], sd[pre + "out_proj.bias"])


# --------------------------------------------------------------------------------------
# model/modules/resampler.py
# --------------------------------------------------------------------------------------


def resampler_forward(x_f, sd, pre="resampler.", heads=8):
    """resampler.py:33-36,46-52 -- x_f [N,B,D] (un-normalised expert tokens) -> [64,B,D]."""
    lat = sd[pre + "latents"].unsqueeze(1).expand(-1, x_f.shape[1], -1)
    l = 0
    while f"{pre}perceiver_blocks.{l}.ln_1.weight" in sd:
        p = f"{pre}perceiver_blocks.{l}."
        ql = layer_norm(lat, sd[p + "ln_1.weight"], sd[p + "ln_1.bias"])
        kv = torch.cat([ql, layer_norm(x_f, sd[p + "ln_2.weight"], sd[p + "ln_2.bias"])], dim=0)
        lat = lat + mha(ql, kv, sd, p + "attn.", heads)
        h = layer_norm(lat, sd[p + "ln_ff.weight"], sd[p + "ln_ff.bias"])
        h = linear(h, sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"])
        h = linear(squared_relu(h), sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"])
        lat = lat + h
        l += 1
    return lat


# --------------------------------------------------------------------------------------
# model/modules/vit.py
# --------------------------------------------------------------------------------------

_STEM_STRIDES = {"seg": (2, 2, 1, 1), "obj_detection": (2, 2, 1, 1), "ocr_detection": (2, 2, 1, 1)}


def stem_forward(x, sd, pre, domain, patch_size, training=False, bn_stats_out=None):
    """vit.py:86-120 -- one modality's conv stem.  ``pre`` = 'conv1.<domain>.'."""
    if domain == "rgb":
        return F.conv2d(x, sd[pre + "weight"], stride=patch_size)
    if domain in _STEM_STRIDES:
        scale, strides = 4.0 / patch_size, _STEM_STRIDES[domain]
    else:
        scale, strides = 16.0 / patch_size, (2, 2, 2, 2)
    if scale != 1.0:
        # nn.UpsamplingBilinear2d == bilinear, align_corners=True
        x = F.interpolate(x, scale_factor=scale, mode="bilinear", align_corners=True)
    for i, s in enumerate(strides):
        ci, bi = 1 + 3 * i, 2 + 3 * i
        x = F.conv2d(x, sd[f"{pre}{ci}.weight"], stride=s, padding=1)
        if training:
            mean = x.mean(dim=(0, 2, 3))
            var = x.var(dim=(0, 2, 3), unbiased=False)
            if bn_stats_out is not None:
                bn_stats_out[f"{pre}{bi}"] = (mean, x.var(dim=(0, 2, 3), unbiased=True))
        else:
            mean, var = sd[f"{pre}{bi}.running_mean"], sd[f"{pre}{bi}.running_var"]
        x = (x - mean[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + 1e-5)
        x = x * sd[f"{pre}{bi}.weight"][None, :, None, None] + sd[f"{pre}{bi}.bias"][None, :, None, None]
        x = torch.relu(x)
    return F.conv2d(x, sd[f"{pre}13.weight"])


def instance_table(instance_map, rng=random):
    """vit.py:144-146 -- one ``random.randint(0,127)`` per unique instance id, ascending
    (includes the background id 255).  Returns {id: row}."""
    return {int(l): rng.randint(0, 127) for l in instance_map.unique().tolist()}


def vit_block(x, sd, p, heads):
    """vit.py:70-75 with ResidualAttentionBlock vit.py:55-59 and Adaptor utils.py:60-65."""
    h = layer_norm(x, sd[p + "0.ln_1.weight"], sd[p + "0.ln_1.bias"])
    x = x + mha(h, h, sd, p + "0.attn.", heads)
    x = adaptor(x, sd, p + "1.", norm_late=False)
    h = layer_norm(x, sd[p + "0.ln_2.weight"], sd[p + "0.ln_2.bias"])
    h = linear(h, sd[p + "0.mlp.c_fc.weight"], sd[p + "0.mlp.c_fc.bias"])
    h = linear(quick_gelu(h), sd[p + "0.mlp.c_proj.weight"], sd[p + "0.mlp.c_proj.bias"])
    return x + h


def encoder_forward(experts: Dict, sd, patch_size: int, heads: Optional[int] = None,
                    training: bool = False, rng=random, tables_out: Optional[dict] = None):
    """VisionTransformer.forward, vit.py:133-172.  ``sd`` holds the encoder's keys
    (no 'expert_encoder.' prefix).  Returns [S, B, D] (seq-first)."""
    D = sd["positional_embedding"].shape[1]
    heads = heads or D // 64
    pos = sd["positional_embedding"]
    rgb_inputs, experts_inputs = None, []
    for exp in experts:
        domain = "seg" if "seg" in exp else exp
        x_ = experts[exp] if exp != "obj_detection" else experts[exp]["label"]
        x_ = stem_forward(x_, sd, f"conv1.{domain}.", domain, patch_size, training)
        if exp == "obj_detection":  # vit.py:141-148
            inst = experts[exp]["instance"]
            imap = F.interpolate(inst.to(x_.dtype), size=x_.shape[2:], mode="nearest")[:, 0]
            table = instance_table(inst, rng)
            if tables_out is not None:
                tables_out.update(table)
            x_ = x_.clone()
            for l, l_ in table.items():
                m = (imap == l)
                x_ = x_ + m[:, None].to(x_.dtype) * sd["instance_embedding"][l_][None, :, None, None]
        x_ = x_.flatten(2).transpose(1, 2)  # b d h w -> b (h w) d
        if domain == "rgb":
            rgb_inputs = x_ + pos
        else:
            experts_inputs.append(x_ + interpolate_pos_embed(pos, x_.shape[1]))
    rgb_inputs = rgb_inputs.transpose(0, 1)
    if experts_inputs:
        xf = torch.cat(experts_inputs, dim=1).transpose(0, 1)
        lat = resampler_forward(xf, sd)
        x = torch.cat([rgb_inputs, lat], dim=0)
    else:
        x = rgb_inputs
    x = layer_norm(x, sd["ln_pre.weight"], sd["ln_pre.bias"])
    l = 0
    while f"transformer.resblocks.{l}.0.ln_1.weight" in sd:
        x = vit_block(x, sd, f"transformer.resblocks.{l}.", heads)
        l += 1
    return layer_norm(x, sd["ln_post.weight"], sd["ln_post.bias"])


# --------------------------------------------------------------------------------------
# model/modules/roberta.py
# --------------------------------------------------------------------------------------


def position_ids(input_ids, padding_idx=1):
    """roberta.py:38-45"""
    mask = input_ids.ne(padding_idx).int()
    return (torch.cumsum(mask, dim=1).type_as(mask) * mask).long() + padding_idx


def embeddings(input_ids, sd, pre="roberta.embeddings."):
    """roberta.py:66-76 (dropout is identity in eval)."""
    e = sd[pre + "word_embeddings.weight"][input_ids] + sd[pre + "token_type_embeddings.weight"][0]
    e = e + sd[pre + "position_embeddings.weight"][position_ids(input_ids)]
    return layer_norm(e, sd[pre + "LayerNorm.weight"], sd[pre + "LayerNorm.bias"])


def extended_mask(attention_mask, dtype=torch.float32):
    """roberta.py:310 -> HF get_extended_attention_mask with config.is_decoder:
    [B,1,T,T] = (1 - causal[i>=j] * pad[j]) * finfo(dtype).min"""
    B, T = attention_mask.shape
    causal = torch.tril(torch.ones(T, T, dtype=dtype))
    m = causal[None, None] * attention_mask[:, None, None, :].to(dtype)
    return (1.0 - m) * torch.finfo(dtype).min


def attn_core(h, kv, sd, p, heads, mask):
    """RobertaSelfAttention.forward, roberta.py:95-126 (eval: dropout identity)."""
    q = linear(h, sd[p + "query.weight"], sd[p + "query.bias"])
    k = linear(kv, sd[p + "key.weight"], sd[p + "key.bias"])
    v = linear(kv, sd[p + "value.weight"], sd[p + "value.bias"])
    B, T, H = q.shape
    S = k.shape[1]
    d = H // heads
    q = q.reshape(B, T, heads, d).transpose(1, 2)
    k = k.reshape(B, S, heads, d).transpose(1, 2)
    v = v.reshape(B, S, heads, d).transpose(1, 2)
    s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(d)
    if mask is not None:
        s = s + mask
        s = torch.max(s, torch.tensor(torch.finfo(s.dtype).min, dtype=s.dtype))
    p_ = torch.softmax(s, dim=-1)
    return torch.matmul(p_, v).transpose(1, 2).reshape(B, T, H)


def attn_block(h, kv, sd, p, heads, mask):
    """RobertaAttention = self + RobertaSelfOutput, roberta.py:129-157."""
    a = attn_core(h, kv, sd, p + "self.", heads, mask)
    a = linear(a, sd[p + "output.dense.weight"], sd[p + "output.dense.bias"])
    return layer_norm(a + h, sd[p + "output.LayerNorm.weight"], sd[p + "output.LayerNorm.bias"])


def mlp_block(h, sd, p):
    """RobertaIntermediate + RobertaOutput, roberta.py:160-183."""
    i = gelu_erf(linear(h, sd[p + "intermediate.dense.weight"], sd[p + "intermediate.dense.bias"]))
    o = linear(i, sd[p + "output.dense.weight"], sd[p + "output.dense.bias"])
    return layer_norm(o + h, sd[p + "output.LayerNorm.weight"], sd[p + "output.LayerNorm.bias"])


def decoder_hidden(input_ids, attention_mask, enc, sd, heads):
    """RobertaModel.forward + RobertaEncoder.forward, roberta.py:223-231,288-334."""
    if attention_mask is None:
        attention_mask = torch.ones_like(input_ids)
    h = embeddings(input_ids, sd)
    h = h.to(enc.dtype) if enc is not None else h
    mask = extended_mask(attention_mask, h.dtype)
    l = 0
    while f"roberta.encoder.layer.{l}.0.attention.self.query.weight" in sd:
        p = f"roberta.encoder.layer.{l}."
        h = attn_block(h, h, sd, p + "0.attention.", heads, mask)
        h = attn_block(h, enc, sd, p + "1.", heads, None)
        h = adaptor(h, sd, p + "2.", norm_late=True)
        h = mlp_block(h, sd, p + "0.")
        l += 1
    p = "roberta.encoder.output_layer."
    h = attn_block(h, h, sd, p + "attention.", heads, mask)
    return mlp_block(h, sd, p)


def lm_head(h, sd, pre="lm_head."):
    """RobertaLMHead.forward, roberta.py:421-426; decoder.weight is tied to word_embeddings
    (roberta.py:352-356) and decoder.bias is lm_head.bias (roberta.py:417-419)."""
    x = gelu_erf(linear(h, sd[pre + "dense.weight"], sd[pre + "dense.bias"]))
    x = layer_norm(x, sd[pre + "layer_norm.weight"], sd[pre + "layer_norm.bias"])
    return linear(x, sd["roberta.embeddings.word_embeddings.weight"], sd[pre + "bias"])


def decoder_forward(input_ids, attention_mask, enc, sd, heads, labels=None):
    """RobertaForCausalLMModified.forward, roberta.py:358-399.  Returns (logits, loss[B] | None)."""
    logits = lm_head(decoder_hidden(input_ids, attention_mask, enc, sd, heads), sd)
    loss = None
    if labels is not None:  # roberta.py:381-387
        sl = logits[:, :-1].contiguous()
        tl = labels[:, 1:].contiguous()
        loss = F.cross_entropy(sl.view(-1, sl.shape[-1]).float(), tl.view(-1), reduction="none",
                               label_smoothing=0.1).view(logits.shape[0], -1).sum(1)
    return logits, loss


# --------------------------------------------------------------------------------------
# model/prismer_caption.py / prismer_vqa.py (pre-tokenised restatement)
# --------------------------------------------------------------------------------------


def split_state_dict(sd):
    enc = {k[len("expert_encoder."):]: v for k, v in sd.items() if k.startswith("expert_encoder.")}
    dec = {k[len("text_decoder."):]: v for k, v in sd.items() if k.startswith("text_decoder.")}
    return enc, dec


def caption_labels(input_ids, prompt_length, pad_id=1):
    """prismer_caption.py:22-26"""
    t = input_ids.masked_fill(input_ids == pad_id, -100)
    if prompt_length > 0:
        t[:, :prompt_length] = -100
    return t


def caption_train_loss(experts, input_ids, attention_mask, prompt_length, sd, patch_size, dec_heads,
                       training_bn=False, rng=random):
    """PrismerCaption.forward(train=True), prismer_caption.py:17-34, on pre-tokenised ids."""
    esd, dsd = split_state_dict(sd)
    enc = encoder_forward(experts, esd, patch_size, training=training_bn, rng=rng).transpose(0, 1)
    labels = caption_labels(input_ids, prompt_length)
    logits, loss = decoder_forward(input_ids, attention_mask, enc, dsd, dec_heads, labels)
    return loss.mean(), logits, enc


def greedy_generate(enc, input_ids, sd, heads, max_length=20, min_length=8, eos=2, pad=1, attention_mask=None):
    """HF greedy search as driven by prismer_caption.py:45-50 / prismer_vqa.py:51-57 with num_beams=1 and the
    no-cache ``prepare_inputs_for_generation`` (roberta.py:401-406): every step re-runs the
    decoder on the full prefix; logits[:, -1] -> MinLength processor (eos=-inf while
    cur_len < min_length) -> argmax; finished rows emit pad; stop at max_length.
    ``attention_mask``: the prompt's mask (right-padded VQA questions, prismer_vqa.py:46-47); HF extends it with ones for
    every generated token, and ``logits[:, -1]`` of a short row is the logits of its last PAD position.
    Pinned against the unmodified reference by tests/golden/prismer_tiny_greedy.npz (oracle/gen_golden_greedy.py)."""
    ids = input_ids.clone()
    mask = torch.ones_like(ids) if attention_mask is None else attention_mask.clone().to(ids.dtype)
    B = ids.shape[0]
    unfinished = torch.ones(B, dtype=torch.long)
    step_logits = []
    while ids.shape[1] < max_length:
        logits, _ = decoder_forward(ids, mask, enc, sd, heads)
        nxt = logits[:, -1].float().clone()
        if ids.shape[1] < min_length:
            nxt[:, eos] = -float("inf")
        step_logits.append(nxt)
        tok = nxt.argmax(dim=-1)
        tok = tok * unfinished + pad * (1 - unfinished)
        ids = torch.cat([ids, tok[:, None]], dim=1)
        mask = torch.cat([mask, torch.ones_like(mask[:, :1])], dim=1)
        unfinished = unfinished * (tok != eos).long()
        if unfinished.max() == 0:
            break
    return ids, step_logits


def rank_answers(enc, start_ids, start_mask, answer_ids, answer_mask, sd, heads, k_test, pad=1):
    """inference == 'rank', prismer_caption.py:59-112 / prismer_vqa.py:64-113."""
    logits, _ = decoder_forward(start_ids, start_mask, enc, sd, heads)
    prob_first = torch.softmax(logits[:, -1], dim=1).index_select(1, answer_ids[:, 0])
    _, topk_ids = prob_first.topk(k_test, dim=1)
    B = enc.shape[0]
    a_ids = torch.cat([answer_ids.index_select(0, t) for t in topk_ids], 0)
    a_att = torch.cat([answer_mask.index_select(0, t) for t in topk_ids], 0)
    tile = lambda x: x.repeat_interleave(k_test, dim=0)  # == reference ``tile`` (prismer_caption.py:115-121)
    ids = torch.cat([tile(start_ids), a_ids], 1).long()
    att = torch.cat([tile(start_mask), a_att], 1)
    targets = ids.masked_fill(ids == pad, -100)
    targets[:, :-answer_ids.shape[1]] = -100
    _, loss = decoder_forward(ids, att, tile(enc), sd, heads, targets)
    lps = (-loss / (targets != -100).sum(-1)).view(-1, k_test)
    mx = lps.argmax(1)
    return topk_ids[torch.arange(B), mx], topk_ids, lps


def beam_generate(enc, input_ids, attention_mask, sd, heads, num_beams, max_length, min_length, length_penalty=1.0,
                  eos=2, pad=1):
    """Beam search as ``text_decoder.generate(num_beams=3, ...)`` runs it for prismer_caption.py:42-50 and
    prismer_vqa.py:45-57 (cache-less decoder pass per step, roberta.py:401-406).

    The search itself is third-party: ``transformers`` (reference pin ~=4.26.1, ``requirements.txt:6``; 5.5.0 in this
    container).  This is a per-sample, list-based restatement of the 5.5.0 procedure -- the version the golden vectors
    in tests/golden/prismer_tiny_beam.npz were produced with (oracle/gen_golden_beam.py):

      every step : log_softmax of the last position -> MinLength (eos = -inf while cur_len < min_length) -> add the
                   running beam scores -> the 2*num_beams best (beam, token) continuations of each sample;
      a continuation "stops" when its token is eos or it reaches max_length;
      running    : the num_beams best continuations that did not stop;
      finished   : stopped continuations ranked inside the first num_beams candidates enter a pool of num_beams finished
                   hypotheses with score sum_logprob / generated_len**length_penalty (generated_len counts the eos);
      early stop : a sample is closed once its pool is full and its best running beam, normalised by the CURRENT generated
                   length, can no longer beat the worst pooled score; the loop ends when every sample is closed or no
                   continuation can be extended.

    Returns (ids [B, L] padded with ``pad``, scores [B])."""
    B, T0 = input_ids.shape
    nb, K = num_beams, 2 * num_beams
    enc_rep = enc.repeat_interleave(nb, dim=0)
    run_seq = [[input_ids[b].tolist() for _ in range(nb)] for b in range(B)]
    run_score = torch.zeros(B, nb)
    run_score[:, 1:] = -1e9
    pool_seq = [[input_ids[b].tolist() for _ in range(nb)] for b in range(B)]
    pool_score = torch.full((B, nb), -1e9)
    pool_done = torch.zeros(B, nb, dtype=torch.bool)
    open_ = [True] * B
    cur = T0
    while True:
        flat = torch.tensor([s for b in range(B) for s in run_seq[b]], dtype=torch.long)
        att = torch.cat([attention_mask.repeat_interleave(nb, dim=0), torch.ones(B * nb, cur - T0, dtype=attention_mask.dtype)], 1)
        logits, _ = decoder_forward(flat, att, enc_rep, sd, heads)
        lp = torch.log_softmax(logits[:, -1].float(), dim=-1)
        V = lp.shape[-1]
        if cur < min_length:
            lp[:, eos] = -float("inf")
        acc = (lp.view(B, nb, V) + run_score[:, :, None]).reshape(B, nb * V)
        top_s, top_i = acc.topk(K, dim=1)
        can_extend = False
        for b in range(B):
            cands = [run_seq[b][int(i) // V] + [int(i) % V] for i in top_i[b]]
            stop = torch.tensor([c[-1] == eos or cur + 1 >= max_length for c in cands])
            can_extend = can_extend or not bool(stop.all())
            # beams that keep running
            alive = top_s[b] + stop.float() * -1.0e9
            keep = alive.topk(nb).indices
            run_seq[b] = [cands[int(j)] for j in keep]
            run_score[b] = alive[keep]
            # pool of finished hypotheses
            fin = top_s[b] / ((cur + 1 - T0) ** length_penalty)
            if not open_[b]:
                fin = fin + -1.0e9
            entered = stop & (torch.arange(K) < nb)
            fin = fin + (~entered).float() * -1.0e9
            m_score = torch.cat([pool_score[b], fin])
            m_seq = pool_seq[b] + cands
            m_done = torch.cat([pool_done[b], entered])
            sel = m_score.topk(nb).indices
            pool_score[b], pool_done[b] = m_score[sel], m_done[sel]
            pool_seq[b] = [m_seq[int(j)] for j in sel]
        cur += 1
        for b in range(B):
            best_running = run_score[b, 0] / ((cur - T0) ** length_penalty)
            worst = pool_score[b].min() if bool(pool_done[b].all()) else torch.tensor(-1.0e9)
            open_[b] = open_[b] and bool(best_running > worst)
        if not (any(open_) and can_extend):
            break
    best = [pool_seq[b][0] for b in range(B)]
    L = max(len(s) for s in best)
    ids = torch.full((B, L), pad, dtype=torch.long)
    for b, s in enumerate(best):
        ids[b, :len(s)] = torch.tensor(s)
    return ids, pool_score[:, 0].clone()


# ---------------------------------------------------------------------------------------------------- model-level glue
def _encode(experts, sd, patch_size, rng):
    esd, dsd = split_state_dict(sd)
    return encoder_forward(experts, esd, patch_size, rng=rng).transpose(0, 1), dsd


def caption_forward(experts, sd, tokenizer, patch_size, heads, caption=None, answer=None, train=True, prefix="",
                    inference="generate", k_test=32, rng=random):
    """PrismerCaption.forward (prismer_caption.py:15-112) on strings; eval-mode modules (no dropout, BN running stats)."""
    enc, dsd = _encode(experts, sd, patch_size, rng)
    B = enc.shape[0]
    if train:                                                                    # :17-34
        tok = tokenizer(caption, padding="longest", truncation=True, max_length=30, return_tensors="pt")
        prompt = len(tokenizer(prefix).input_ids) - 1 if len(prefix) > 0 else 0
        labels = caption_labels(tok.input_ids, prompt, tokenizer.pad_token_id)
        _, loss = decoder_forward(tok.input_ids, tok.attention_mask, enc, dsd, heads, labels)
        return loss.mean()
    ptok = tokenizer([prefix] * B, padding="longest", return_tensors="pt")
    start_ids, start_mask = ptok.input_ids[:, :-1], ptok.attention_mask[:, :-1]  # :38-40, :70-71 (drop </s>)
    if inference == "generate":                                                  # :36-57
        ids, _ = beam_generate(enc, start_ids, start_mask, dsd, heads, 3, 20, 8)
        cut = len(prefix) + (1 if len(prefix) > 0 else 0)
        return [tokenizer.decode(row, skip_special_tokens=True)[cut:] for row in ids]
    atok = tokenizer([" " + a.lower() + "</s>" for a in answer], padding="longest", return_tensors="pt", add_special_tokens=False)
    return rank_answers(enc, start_ids, start_mask, atok.input_ids, atok.attention_mask, dsd, heads, k_test, tokenizer.pad_token_id)[0]


def vqa_forward(experts, sd, tokenizer, patch_size, heads, question, answer=None, weights=None, train=True, inference="rank",
                k_test=128, rng=random):
    """PrismerVQA.forward (prismer_vqa.py:16-113) on strings; eval-mode modules."""
    enc, dsd = _encode(experts, sd, patch_size, rng)
    q = tokenizer(["<s>" + s.capitalize() for s in question], padding="longest", truncation=True, max_length=35,
                  add_special_tokens=False, return_tensors="pt")                   # :18-20
    if not train and inference == "generate":                                     # :44-62
        T0 = q.input_ids.shape[1]
        ids, _ = beam_generate(enc, q.input_ids, q.attention_mask, dsd, heads, 3, T0 + 10, T0 + 2, length_penalty=-1)
        return [tokenizer.decode(row[T0:], skip_special_tokens=True).lower().strip() for row in ids]
    a = tokenizer([" " + s.capitalize() + "</s>" for s in answer], padding="longest", return_tensors="pt", add_special_tokens=False)
    if train:                                                                     # :22-42
        ids = torch.cat([q.input_ids, a.input_ids], 1).long()
        att = torch.cat([q.attention_mask, a.attention_mask], 1)
        targets = ids.masked_fill(ids == tokenizer.pad_token_id, -100)
        targets[:, :-a.input_ids.shape[1]] = -100
        _, loss = decoder_forward(ids, att, enc, dsd, heads, targets)
        return (weights * loss).mean()
    return rank_answers(enc, q.input_ids, q.attention_mask, a.input_ids, a.attention_mask, dsd, heads, k_test, tokenizer.pad_token_id)[0]


# ---------------------------------------------------------------------------------------------------- expert-label post-processing
def post_label_process(inputs, labels_info, features, eps=1e-6):
    """dataset/utils.py:117-160 restated with numpy-style indexing.  ``inputs``: what ``Transform`` returns (:56-63) --
    float [c,H,W] in [0,1] for depth / normal / edge, int64 [1,H,W] label maps otherwise.  ``features``: the tables the
    reference loads as module globals (:17-20): {'coco','ade','detection': [N,64], 'background': [64]}."""
    out = {}
    for exp, x in inputs.items():
        if exp in ("depth", "normal", "edge"):                       # :120-121  remap to [-1, 1] with the per-sample range
            out[exp] = 2 * (x - x.min()) / (x.max() - x.min() + eps) - 1
            continue
        if exp == "rgb":
            out[exp] = x
            continue
        ids = x[0]
        emb = torch.empty((64, *ids.shape), dtype=torch.float32)
        for l in torch.unique(ids).tolist():
            if l == 255:
                row = features["background"]                          # :126,135,145,155
            elif exp == "seg_coco":
                row = features["coco"][l]                             # :128
            elif exp == "seg_ade":
                row = features["ade"][l]                              # :137
            elif exp == "obj_detection":
                row = features["detection"][labels_info[exp][str(l)]]  # :147 (json keys are strings)
            elif exp == "ocr_detection":
                row = labels_info[exp][l]["features"]                 # :157
            else:
                raise KeyError(exp)
            emb[:, ids == l] = row[:, None]
        out[exp] = {"label": emb, "instance": x} if exp == "obj_detection" else emb   # :148
    return out
