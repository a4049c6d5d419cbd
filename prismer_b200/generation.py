"""Decoding loops for ``RobertaForCausalLMModified.generate`` (prismer_caption.py:45-50, prismer_vqa.py:51-57).

Greedy follows HF greedy search as the reference drives it: ``logits[:, -1]`` -> MinLength processor (eos = -inf while
cur_len < min_length) -> argmax (device kernel, lowest index on ties) -> finished rows emit pad -> stop at max_length
or when every row has produced eos.  The per-step decoder pass re-runs on the full prefix exactly like the reference's
cache-less ``prepare_inputs_for_generation`` (roberta.py:401-406); the cross-attention K/V projections of the visual
tokens are computed once per call instead of once per step and layer."""
from __future__ import annotations

import torch

from . import engine, ops


@torch.no_grad()
def greedy(dec, input_ids, enc, attention_mask, max_length=20, min_length=0, return_step_logits=False):
    cfg = dec.config
    eos, pad, V = cfg.eos_token_id, cfg.pad_token_id, cfg.vocab_size
    engine._store(dec).refresh()
    B, T0 = input_ids.shape
    dev = input_ids.device
    ids = torch.full((B, max_length), pad, dtype=torch.int64, device=dev)
    ids[:, :T0] = input_ids
    unfinished = torch.ones(B, dtype=torch.int64, device=dev)
    cur = T0
    steps = []
    kv = engine.cross_kv(dec, enc)              # visual K/V of all layers: once per call
    ones = torch.ones((B, max_length), dtype=torch.int64, device=dev)
    while cur < max_length:
        cur_ids = ids[:, :cur].contiguous()
        last, _, _, _ = engine.decoder_forward(dec, cur_ids, ones[:, :cur].contiguous(), enc, None, None, save=False, kv=kv,
                                               last_only=True)      # [B, V] fp32 logits of the last position
        tok = ops.argmax(last, V, suppress_eos=cur < min_length, eos=eos)
        if return_step_logits:
            steps.append(last.clone())
        tok = tok * unfinished + pad * (1 - unfinished)
        ids[:, cur] = tok
        unfinished = unfinished * (tok != eos).long()
        cur += 1
        if int(unfinished.max()) == 0:
            break
    out = ids[:, :cur]
    return (out, steps) if return_step_logits else out


@torch.no_grad()
def beam_search(dec, input_ids, enc, attention_mask, num_beams, max_length, min_length, length_penalty=1.0):
    """HF-style beam search (2*num_beams candidates, length-normalised scores).  Bookkeeping on the host; the decoder
    passes and the log-softmax statistics run on the device."""
    cfg = dec.config
    eos, pad, V = cfg.eos_token_id, cfg.pad_token_id, cfg.vocab_size
    engine._store(dec).refresh()
    B, T0 = input_ids.shape
    dev = input_ids.device
    nb = num_beams
    enc_b = enc.repeat_interleave(nb, dim=0) if enc.is_contiguous() else enc.contiguous().repeat_interleave(nb, dim=0)
    seqs = input_ids.repeat_interleave(nb, dim=0)
    beam_scores = torch.zeros((B, nb), dtype=torch.float32, device=dev)
    beam_scores[:, 1:] = -1e9
    beam_scores = beam_scores.view(-1)
    done = [False] * B
    hyps = [[] for _ in range(B)]      # (score, tensor)
    cur = T0
    while cur < max_length:
        logits, _, _, _ = engine.decoder_forward(dec, seqs.contiguous(), torch.ones_like(seqs), enc_b, None, None, save=False)
        last = logits.view(B * nb, cur, -1)[:, -1].float()
        lp = last - torch.logsumexp(last, dim=-1, keepdim=True)
        if cur < min_length:
            lp[:, eos] = -float("inf")
        scores = (lp + beam_scores[:, None]).view(B, nb * V)
        top_s, top_i = scores.topk(2 * nb, dim=1)
        top_s, top_i = top_s.cpu(), top_i.cpu()
        new_seqs, new_scores = [], []
        for b in range(B):
            cand = []
            for s, i in zip(top_s[b].tolist(), top_i[b].tolist()):
                beam, tok = i // V, i % V
                if tok == eos:
                    if len(cand) < nb and not done[b]:
                        hyps[b].append((s / ((cur + 1 - 0) ** length_penalty), seqs[b * nb + beam].clone()))
                    continue
                cand.append((s, beam, tok))
                if len(cand) == nb:
                    break
            if len(hyps[b]) >= nb:
                best_possible = cand[0][0] / ((cur + 1) ** length_penalty) if cand else -1e30
                worst = sorted(h[0] for h in hyps[b])[-nb]
                done[b] = done[b] or worst >= best_possible
            for s, beam, tok in cand:
                new_seqs.append(torch.cat([seqs[b * nb + beam], torch.tensor([tok], device=dev)]))
                new_scores.append(s)
        seqs = torch.stack(new_seqs)
        beam_scores = torch.tensor(new_scores, dtype=torch.float32, device=dev)
        cur += 1
        if all(done):
            break
    out = []
    for b in range(B):
        if not done[b] or len(hyps[b]) < 1:
            for k in range(nb):
                hyps[b].append((float(beam_scores[b * nb + k]) / (cur ** length_penalty), seqs[b * nb + k]))
        best = max(hyps[b], key=lambda h: h[0])[1]
        out.append(best)
    L = max(len(o) for o in out)
    res = torch.full((B, L), pad, dtype=torch.int64, device=dev)
    for b, o in enumerate(out):
        res[b, :len(o)] = o
    return res
