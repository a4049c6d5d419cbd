"""Decoding loops for ``RobertaForCausalLMModified.generate`` (prismer_caption.py:45-50, prismer_vqa.py:51-57).

Greedy follows HF greedy search as the reference drives it: ``logits[:, -1]`` -> MinLength processor (eos = -inf while
cur_len < min_length) -> argmax (device kernel, lowest index on ties) -> finished rows emit pad -> stop at max_length
or when every row has produced eos.  The per-step decoder pass re-runs on the full prefix exactly like the reference's
cache-less ``prepare_inputs_for_generation`` (roberta.py:401-406); the cross-attention K/V projections of the visual
tokens are computed once per call instead of once per step and layer."""
from __future__ import annotations

import torch

from . import engine, ops


def _greedy_loop(dec, ids, T0, enc, max_length, min_length, early_exit, steps=None):
    """Device-side greedy loop on a preallocated ``ids`` [B, max_length] buffer (prefix already in columns [0, T0)).
    With ``early_exit=False`` there is no host synchronisation at all (finished rows keep emitting pad), so the whole loop
    can be captured in a CUDA graph."""
    cfg = dec.config
    eos, pad, V = cfg.eos_token_id, cfg.pad_token_id, cfg.vocab_size
    B = ids.shape[0]
    dev = ids.device
    unfinished = torch.ones(B, dtype=torch.int64, device=dev)
    kv = engine.cross_kv(dec, enc)              # visual K/V of all layers: once per call, not once per step and layer
    ones = torch.ones((B, max_length), dtype=torch.int64, device=dev)
    cur = T0
    while cur < max_length:
        cur_ids = ids[:, :cur].contiguous()
        last, _, _, _ = engine.decoder_forward(dec, cur_ids, ones[:, :cur].contiguous(), enc, None, None, save=False, kv=kv,
                                               last_only=True)      # [B, V] fp32 logits of the last position
        tok = ops.argmax(last, V, suppress_eos=cur < min_length, eos=eos)
        if steps is not None:
            steps.append(last.clone())
        tok = tok * unfinished + pad * (1 - unfinished)
        ids[:, cur] = tok
        unfinished = unfinished * (tok != eos).long()
        cur += 1
        if early_exit and int(unfinished.max()) == 0:
            break
    return cur


@torch.no_grad()
def greedy(dec, input_ids, enc, attention_mask, max_length=20, min_length=0, return_step_logits=False):
    engine._store(dec).refresh()
    B, T0 = input_ids.shape
    ids = torch.full((B, max_length), dec.config.pad_token_id, dtype=torch.int64, device=input_ids.device)
    ids[:, :T0] = input_ids
    steps = [] if return_step_logits else None
    cur = _greedy_loop(dec, ids, T0, enc, max_length, min_length, early_exit=True, steps=steps)
    out = ids[:, :cur]
    return (out, steps) if return_step_logits else out


def trim_finished(ids: torch.Tensor, T0: int, eos: int) -> torch.Tensor:
    """HF stops as soon as every row has produced eos: cut the fixed-length output of the graphed loop at that column."""
    gen = ids[:, T0:]
    hit = (gen == eos)
    if bool(hit.any(dim=1).all()):
        last = int(hit.float().argmax(dim=1).max()) + 1
        return ids[:, :T0 + last]
    return ids


class GraphedCaptioner:
    """Encoder forward + the complete greedy decode (prismer_caption.py:36-50 with num_beams=1) captured in ONE CUDA graph
    for a fixed (batch, prefix length, max_length): ~3400 kernel launches per batch become one ``cudaGraphLaunch``.
    The instance-embedding table (host ``random.randint``, vit.py:144-146) is drawn before every replay."""

    def __init__(self, model, experts, prefix_ids, max_length=20, min_length=8):
        self.model = model
        vit, dec = model.expert_encoder, model.text_decoder
        st = self.store = engine._store(model)
        st.refresh()
        dev = st.device
        self.experts = {k: ({kk: vv.clone() for kk, vv in v.items()} if isinstance(v, dict) else v.clone())
                        for k, v in engine._canon_experts(experts).items()}
        self.T0, self.max_length, self.min_length = prefix_ids.shape[1], max_length, min_length
        self.prefix = prefix_ids.clone()
        self.ids = torch.full((prefix_ids.shape[0], max_length), dec.config.pad_token_id, dtype=torch.int64, device=dev)
        self.has_inst = "obj_detection" in self.experts
        self.table = torch.zeros(256, dtype=torch.int32, device=dev) if self.has_inst else None

        def run():
            self.ids.fill_(dec.config.pad_token_id)
            self.ids[:, :self.T0] = self.prefix
            out, S, B, _ = engine.encoder_forward(vit, self.experts, save=False, inst_table=self.table)
            enc = out.view(S, B, -1).transpose(0, 1)
            _greedy_loop(dec, self.ids, self.T0, enc, max_length, min_length, early_exit=False)

        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s), torch.no_grad():
            self._draw_table()
            run()
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            run()

    def _draw_table(self):
        if self.has_inst:
            self.table.copy_(engine._instance_table(self.experts["obj_detection"]["instance"]), non_blocking=True)

    def load_inputs(self, experts, non_blocking=True):
        for k, v in experts.items():
            if isinstance(v, dict):
                for kk, vv in v.items():
                    self.experts[k][kk].copy_(vv, non_blocking=non_blocking)
            else:
                self.experts[k].copy_(v, non_blocking=non_blocking)

    def __call__(self) -> torch.Tensor:
        """Replay on the current static inputs; returns the [B, max_length] id buffer (use ``trim_finished`` for HF's length)."""
        self.store.refresh()
        self._draw_table()
        self.graph.replay()
        return self.ids


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
