"""Decoding loops for ``RobertaForCausalLMModified.generate`` (prismer_caption.py:45-50, prismer_vqa.py:51-57).

Greedy follows HF greedy search as the reference drives it: ``logits[:, -1]`` -> MinLength processor (eos = -inf while
cur_len < min_length) -> argmax (device kernel, lowest index on ties) -> finished rows emit pad -> stop at max_length
or when every row has produced eos.  Two schedules compute the per-step logits:

* ``KV_CACHE = True`` (default): one new token per sequence and step against per-layer K/V caches (``kv_decode.py``,
  ``csrc/decode.cu``) -- the algorithmic 13.6 GFLOP/img instead of the reference's 178;
* ``KV_CACHE = False``: the decoder re-runs on the full prefix every step exactly like the reference's cache-less
  ``prepare_inputs_for_generation`` (roberta.py:401-406) -- kept as the in-repo cross-check of the cached path (and used by beam search).

In both, the cross-attention K/V projections of the visual tokens are computed once per call instead of once per step and layer."""
from __future__ import annotations

import torch

from . import engine, kv_decode, ops

KV_CACHE = True


def _greedy_loop(dec, ids, T0, enc, max_length, min_length, early_exit, steps=None, prompt_mask=None):
    cfg = dec.config
    if KV_CACHE and cfg.hidden_size // cfg.num_attention_heads == 64 and enc.shape[1] <= 320:
        return kv_decode.greedy_loop(dec, ids, T0, enc, max_length, min_length, early_exit, steps, prompt_mask)
    return _greedy_loop_nocache(dec, ids, T0, enc, max_length, min_length, early_exit, steps, prompt_mask)


def _greedy_loop_nocache(dec, ids, T0, enc, max_length, min_length, early_exit, steps=None, prompt_mask=None):
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
    if prompt_mask is not None:                 # right-padded VQA questions (prismer_vqa.py:46-47); generated tokens attend
        ones[:, :T0] = prompt_mask.to(torch.int64)
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
    cur = _greedy_loop(dec, ids, T0, enc, max_length, min_length, early_exit=True, steps=steps, prompt_mask=attention_mask)
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
        self.experts = engine.clone_experts(engine._canon_experts(experts))
        self.T0, self.max_length, self.min_length = prefix_ids.shape[1], max_length, min_length
        self.prefix = prefix_ids.clone()
        self.ids = torch.full((prefix_ids.shape[0], max_length), dec.config.pad_token_id, dtype=torch.int64, device=dev)
        self.has_inst = "obj_detection" in self.experts
        self.table = torch.zeros(256, dtype=torch.int32, device=dev) if self.has_inst else None
        self.presence = engine.InstancePresence(dev).request(engine.instance_map(self.experts["obj_detection"])) if self.has_inst else None

        def run():
            self.ids.fill_(dec.config.pad_token_id)
            self.ids[:, :self.T0] = self.prefix
            out, S, B, _ = engine.encoder_forward(vit, self.experts, save=False, inst_table=self.table)
            self.S = S
            enc = out.view(S, B, -1).transpose(0, 1)
            _greedy_loop(dec, self.ids, self.T0, enc, max_length, min_length, early_exit=False)

        from . import _C
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s), torch.no_grad():
            self._draw_table()
            c0 = _C.CALLS
            run()
            self.launches = _C.CALLS - c0          # C-ABI calls (>= 1 kernel each) captured in the graph
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            run()

    def _draw_table(self):
        if self.has_inst:       # flags were computed when the inputs were loaded (engine.InstancePresence): no blocking D2H here
            self.table.copy_(self.presence.table())     # 1 KiB from pageable memory: staged by the driver at call time, stream-ordered

    def load_inputs(self, experts, non_blocking=True, presence=None):
        engine.copy_experts_(self.experts, experts, non_blocking)
        if self.has_inst:
            self.presence = presence if presence is not None else engine.InstancePresence(self.store.device).request(
                engine.instance_map(self.experts["obj_detection"]))

    def __call__(self) -> torch.Tensor:
        """Replay on the current static inputs; returns the [B, max_length] id buffer (use ``trim_finished`` for HF's length)."""
        self.store.refresh()
        self._draw_table()
        self.graph.replay()
        return self.ids


def beam_search_core(step_logits, input_ids, attention_mask, num_beams, max_length, min_length, length_penalty, eos, pad):
    """Batched beam search with all bookkeeping in fixed-shape device tensors (no per-hypothesis host objects): the
    procedure ``text_decoder.generate(num_beams=3, ...)`` runs for prismer_caption.py:42-50 / prismer_vqa.py:45-57, i.e.
    the ``transformers`` beam search (reference pin ~=4.26.1; the 5.5.0 procedure is the one that could be run and pinned
    here -- tests/golden/prismer_tiny_beam.npz, oracle/gen_golden_beam.py).

    ``step_logits(ids [B*nb, cur], mask [B*nb, cur]) -> [B*nb, V]`` fp32 logits of the last position.

    State per sample: ``run_*`` the num_beams live beams, ``pool_*`` the num_beams best finished hypotheses (score =
    sum_logprob / generated_len**length_penalty, eos counted), ``open_`` whether a live beam can still beat the pool.
    One scalar is read back per step (loop exit test); everything else stays on the device.
    Returns (ids [B, L] padded with ``pad``, scores [B])."""
    B, T0 = input_ids.shape
    dev = input_ids.device
    nb, K = num_beams, 2 * num_beams
    f32 = torch.float32
    run_seq = torch.full((B, nb, max_length), pad, dtype=torch.int64, device=dev)
    run_seq[:, :, :T0] = input_ids[:, None, :]
    pool_seq = run_seq.clone()
    run_score = torch.zeros((B, nb), dtype=f32, device=dev)
    run_score[:, 1:] = -1.0e9
    pool_score = torch.full((B, nb), -1.0e9, dtype=f32, device=dev)
    pool_done = torch.zeros((B, nb), dtype=torch.bool, device=dev)
    pool_len = torch.full((B, nb), T0, dtype=torch.int64, device=dev)
    open_ = torch.ones((B, 1), dtype=torch.bool, device=dev)
    in_top = (torch.arange(K, device=dev) < nb)[None, :]
    mask = torch.ones((B * nb, max_length), dtype=torch.int64, device=dev)
    if attention_mask is not None:
        mask[:, :T0] = attention_mask.to(torch.int64).repeat_interleave(nb, dim=0)
    take = lambda t, idx: t.gather(1, idx[:, :, None].expand(-1, -1, t.shape[2]))
    cur = T0
    while cur < max_length:
        logits = step_logits(run_seq[:, :, :cur].reshape(B * nb, cur), mask[:, :cur])
        lp = torch.log_softmax(logits.to(f32), dim=-1)
        V = lp.shape[-1]
        if cur < min_length:
            lp[:, eos] = -float("inf")                      # MinLengthLogitsProcessor
        acc = (lp.reshape(B, nb, V) + run_score[:, :, None]).reshape(B, nb * V)     # (logits may be a strided view of a padded buffer)
        top_s, top_i = acc.topk(K, dim=1)                   # 2*nb continuations: enough live ones survive nb eos hits
        cand = take(run_seq, top_i // V)
        tok = top_i % V
        cand[:, :, cur] = tok
        stop = (tok == eos) | (cur + 1 >= max_length)
        # live beams of the next step
        alive = top_s + stop.to(f32) * -1.0e9
        keep = alive.topk(nb, dim=1).indices
        run_seq, run_score = take(cand, keep), alive.gather(1, keep)
        # finished pool: only stopped continuations ranked inside the first nb, and only while the sample is open
        entered = stop & in_top
        fin = top_s / ((cur + 1 - T0) ** length_penalty)
        fin = fin + (~open_).to(f32) * -1.0e9
        fin = fin + (~entered).to(f32) * -1.0e9
        m_score = torch.cat([pool_score, fin], dim=1)
        sel = m_score.topk(nb, dim=1).indices
        pool_score = m_score.gather(1, sel)
        pool_seq = take(torch.cat([pool_seq, cand], dim=1), sel)
        pool_done = torch.cat([pool_done, entered], dim=1).gather(1, sel)
        pool_len = torch.cat([pool_len, torch.full((B, K), cur + 1, dtype=torch.int64, device=dev)], dim=1).gather(1, sel)
        cur += 1
        # can the best live beam (normalised by the current generated length) still beat the worst pooled hypothesis?
        best_running = run_score[:, :1] / ((cur - T0) ** length_penalty)
        worst = torch.where(pool_done, pool_score.min(dim=1, keepdim=True).values, torch.full_like(pool_score, -1.0e9))
        open_ = open_ & (best_running > worst).any(dim=1, keepdim=True)
        if not bool(open_.any() & ~stop.all()):
            break
    L = int(pool_len[:, 0].max())
    return pool_seq[:, 0, :L].contiguous(), pool_score[:, 0].contiguous()


@torch.no_grad()
def beam_search(dec, input_ids, enc, attention_mask, num_beams, max_length, min_length, length_penalty=1.0, return_scores=False):
    """``generate(num_beams>1)``: ``beam_search_core`` driven by the CUDA decoder.  Like the reference's cache-less
    ``prepare_inputs_for_generation`` (roberta.py:401-406) every step re-runs the decoder on the full prefix; the
    cross-attention K/V of the (beam-expanded, prismer_caption.py:45 via ``_expand_inputs_for_generation``) visual tokens are
    projected once per call and only the last position goes through the LM head."""
    cfg = dec.config
    engine._store(dec).refresh()
    enc_b = enc.contiguous().repeat_interleave(num_beams, dim=0)
    kv = engine.cross_kv(dec, enc_b)

    def step_logits(ids, mask):
        last, _, _, _ = engine.decoder_forward(dec, ids.contiguous(), mask.contiguous(), enc_b, None, None, save=False, kv=kv,
                                               last_only=True)
        return last

    ids, scores = beam_search_core(step_logits, input_ids, attention_mask, num_beams, max_length, min_length, length_penalty,
                                   cfg.eos_token_id, cfg.pad_token_id)
    return (ids, scores) if return_scores else ids
