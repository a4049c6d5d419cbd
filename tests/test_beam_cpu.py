"""Beam search (prismer_caption.py:42-50, prismer_vqa.py:45-57) without a GPU:

  * the oracle restatement ``oracle.prismer_oracle.beam_generate`` is pinned against golden ids + sequence scores that
    ``oracle/gen_golden_beam.py`` produced with the UNMODIFIED reference decoder under ``transformers.generate``;
  * the product's fixed-shape bookkeeping ``prismer_b200.generation.beam_search_core`` is run on CPU tensors with the oracle
    decoder supplying the per-step logits, and must return the same ids bit-exactly -- so on the GPU only the logits differ.
"""
import numpy as np
import pytest
import torch

from oracle import prismer_oracle as O
from prismer_b200 import generation
from tests.helpers import (BEAM_CASES, TINY_DEC, beam_case_inputs, beam_decoder_state, decoder_template, load_beam_golden)

HEADS = TINY_DEC["num_attention_heads"]


@pytest.fixture(scope="module")
def gold():
    return load_beam_golden()


@pytest.fixture(scope="module")
def template():
    return decoder_template()


def _case(c, template, gold):
    sd = beam_decoder_state(template, c["boost"])
    ids, mask, enc = beam_case_inputs(c)
    assert np.array_equal(ids.numpy(), gold[c["name"] + ".prompt"]) and np.array_equal(mask.numpy(), gold[c["name"] + ".mask"])
    return sd, ids, mask, enc


@pytest.mark.parametrize("c", BEAM_CASES, ids=[c["name"] for c in BEAM_CASES])
def test_oracle_beam_matches_reference(c, template, gold):
    sd, ids, mask, enc = _case(c, template, gold)
    with torch.no_grad():
        out, sc = O.beam_generate(enc, ids, mask, sd, HEADS, c["nb"], c["T0"] + c["max_add"], c["T0"] + c["min_add"], c["lp"])
    assert np.array_equal(out.numpy(), gold[c["name"] + ".ids"])            # token ids: bit-exact
    np.testing.assert_allclose(sc.numpy(), gold[c["name"] + ".scores"], rtol=2e-6, atol=2e-5)


@pytest.mark.parametrize("c", BEAM_CASES, ids=[c["name"] for c in BEAM_CASES])
def test_product_bookkeeping_matches_reference(c, template, gold):
    sd, ids, mask, enc = _case(c, template, gold)
    enc_b = enc.repeat_interleave(c["nb"], dim=0)

    def step_logits(flat_ids, flat_mask):
        logits, _ = O.decoder_forward(flat_ids, flat_mask, enc_b, sd, HEADS)
        return logits[:, -1].float()

    with torch.no_grad():
        out, sc = generation.beam_search_core(step_logits, ids, mask, c["nb"], c["T0"] + c["max_add"], c["T0"] + c["min_add"],
                                              c["lp"], TINY_DEC["eos_token_id"], TINY_DEC["pad_token_id"])
    assert np.array_equal(out.numpy(), gold[c["name"] + ".ids"])
    np.testing.assert_allclose(sc.numpy(), gold[c["name"] + ".scores"], rtol=2e-6, atol=2e-5)


def test_beam_core_without_prompt_mask_equals_all_ones(template, gold):
    c = BEAM_CASES[1]
    sd, ids, mask, enc = _case(c, template, gold)
    enc_b = enc.repeat_interleave(c["nb"], dim=0)
    step = lambda i, m: O.decoder_forward(i, m, enc_b, sd, HEADS)[0][:, -1].float()
    with torch.no_grad():
        out, _ = generation.beam_search_core(step, ids, None, c["nb"], c["T0"] + c["max_add"], c["T0"] + c["min_add"], c["lp"], 2, 1)
    assert np.array_equal(out.numpy(), gold[c["name"] + ".ids"])


def test_beam_search_driver_plumbing_with_a_stand_in_engine(template, gold, monkeypatch):
    """``generation.beam_search`` itself (beam expansion of the visual tokens, one K/V projection per call, last-position logits,
    prompt mask) with the engine's three entry points replaced by oracle-backed stand-ins: same ids as the reference."""
    from types import SimpleNamespace
    from prismer_b200 import engine
    c = BEAM_CASES[5]                                     # VQA-style: ragged prompts, length_penalty -1
    sd, ids, mask, enc = _case(c, template, gold)
    calls = {"kv": 0, "fwd": 0}

    def fake_cross_kv(dec, enc_b):
        calls["kv"] += 1
        return SimpleNamespace(B=enc_b.shape[0], enc=enc_b)

    def fake_decoder_forward(dec, input_ids, attention_mask, enc_b, labels, weights, save, kv=None, last_only=False, **_):
        calls["fwd"] += 1
        assert kv is not None and last_only and not save and kv.B == input_ids.shape[0] == attention_mask.shape[0]
        logits, _ = O.decoder_forward(input_ids, attention_mask, kv.enc.float(), sd, HEADS)
        return logits[:, -1].float(), None, None, None

    monkeypatch.setattr(engine, "cross_kv", fake_cross_kv)
    monkeypatch.setattr(engine, "decoder_forward", fake_decoder_forward)
    monkeypatch.setattr(engine, "_store", lambda m: SimpleNamespace(refresh=lambda: None))
    dec = SimpleNamespace(config=SimpleNamespace(eos_token_id=TINY_DEC["eos_token_id"], pad_token_id=TINY_DEC["pad_token_id"],
                                                 vocab_size=TINY_DEC["vocab_size"]))
    out, sc = generation.beam_search(dec, ids, enc, mask, c["nb"], c["T0"] + c["max_add"], c["T0"] + c["min_add"], c["lp"], return_scores=True)
    assert calls["kv"] == 1 and calls["fwd"] >= 2
    assert np.array_equal(out.numpy(), gold[c["name"] + ".ids"])
    np.testing.assert_allclose(sc.numpy(), gold[c["name"] + ".scores"], rtol=2e-6, atol=2e-5)


# ---------------------------------------------------------------------------------------------------- greedy, right-padded prompts
from tests.helpers import GREEDY_CASES, load_greedy_golden  # noqa: E402


@pytest.mark.parametrize("c", GREEDY_CASES, ids=[c["name"] for c in GREEDY_CASES])
def test_oracle_greedy_matches_reference_on_padded_prompts(c, template):
    """HF greedy of the unmodified reference decoder (oracle/gen_golden_greedy.py) vs the oracle restatement: ids bit-exact, incl.
    rows whose prompt is right-padded (first generated token comes from the PAD position's logits)."""
    gold = load_greedy_golden()
    sd = beam_decoder_state(template, c["boost"])
    ids, mask, enc = beam_case_inputs(c)
    assert np.array_equal(ids.numpy(), gold[c["name"] + ".prompt"]) and np.array_equal(mask.numpy(), gold[c["name"] + ".mask"])
    with torch.no_grad():
        out, _ = O.greedy_generate(enc, ids, sd, HEADS, c["T0"] + c["max_add"], c["T0"] + c["min_add"], attention_mask=mask)
    assert np.array_equal(out.numpy(), gold[c["name"] + ".ids"])


@pytest.mark.parametrize("c", GREEDY_CASES, ids=[c["name"] for c in GREEDY_CASES])
def test_product_greedy_loop_matches_reference_with_a_stand_in_engine(c, template, monkeypatch):
    """``generation.greedy`` (prompt mask extended with ones, MinLength, pad after eos, early exit) with the engine's entry points
    replaced by oracle-backed stand-ins on CPU tensors: the same ids as the reference's HF greedy."""
    from types import SimpleNamespace
    from prismer_b200 import engine, ops
    gold = load_greedy_golden()
    sd = beam_decoder_state(template, c["boost"])
    ids, mask, enc = beam_case_inputs(c)

    def fake_decoder_forward(dec, input_ids, attention_mask, enc_b, labels, weights, save, kv=None, last_only=False, **_):
        assert last_only and not save
        logits, _ = O.decoder_forward(input_ids, attention_mask, kv.enc.float(), sd, HEADS)
        return logits[:, -1].float(), None, None, None

    def fake_argmax(logits, V, suppress_eos=False, eos=2):
        l = logits[:, :V].clone()
        if suppress_eos:
            l[:, eos] = -float("inf")
        return l.argmax(dim=-1)

    monkeypatch.setattr(engine, "cross_kv", lambda dec, e: SimpleNamespace(B=e.shape[0], enc=e))
    monkeypatch.setattr(engine, "decoder_forward", fake_decoder_forward)
    monkeypatch.setattr(engine, "_store", lambda m: SimpleNamespace(refresh=lambda: None))
    monkeypatch.setattr(ops, "argmax", fake_argmax)
    monkeypatch.setattr(generation, "KV_CACHE", False)       # the cache-less loop is the one that calls decoder_forward per step
    dec = SimpleNamespace(config=SimpleNamespace(eos_token_id=TINY_DEC["eos_token_id"], pad_token_id=TINY_DEC["pad_token_id"],
                                                 vocab_size=TINY_DEC["vocab_size"]))
    out = generation.greedy(dec, ids, enc, mask, max_length=c["T0"] + c["max_add"], min_length=c["T0"] + c["min_add"])
    assert np.array_equal(out.numpy(), gold[c["name"] + ".ids"])
