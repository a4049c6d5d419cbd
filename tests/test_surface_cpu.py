"""Pins the oracle's model-level glue (``oracle.prismer_oracle.caption_forward / vqa_forward``: tokenisation, label and prompt
masking, answer weights, rank, beam-3 generate + string decoding) against values produced by the reference's OWN
``PrismerCaption.forward`` / ``PrismerVQA.forward`` (oracle/gen_golden_surface.py -> tests/golden/prismer_tiny_surface.npz)."""
import random

import numpy as np
import pytest
import torch

from oracle import prismer_oracle as O
from prismer_b200 import synthetic
from prismer_b200.modeling import template_state_dict
from prismer_b200.tokenizer import HashTokenizer
from tests.helpers import GOLD, SURFACE as S, TINY_DEC

FULL = synthetic.DEFAULT_EXPERTS
HEADS = TINY_DEC["num_attention_heads"]


@pytest.fixture(scope="module")
def fx():
    cfg = S["cfg"]
    tmpl = template_state_dict(width=cfg["width"], layers=cfg["layers"], patch=cfg["patch"], res=cfg["res"], experts=FULL, dec_cfg=TINY_DEC)
    sd = synthetic.synth_state_dict(tmpl, cfg["seed"])
    ex = synthetic.synth_experts(cfg["B"], cfg["res"], FULL, cfg["label"], cfg["in_seed"])
    gold = dict(np.load(f"{GOLD}/prismer_tiny_surface.npz"))
    return cfg, sd, ex, HashTokenizer(TINY_DEC["vocab_size"]), gold


def _run(fx, fn, *a, **k):
    cfg, sd, ex, tok, _ = fx
    random.seed(cfg["py_seed"])
    with torch.no_grad():
        return fn(ex, sd, tok, cfg["patch"], HEADS, *a, **k)


def test_caption_train_loss(fx):
    g = fx[-1]
    assert abs(float(_run(fx, O.caption_forward, caption=S["captions"], prefix=S["prefix"])) - float(g["cap.loss"])) < 2e-5 * float(g["cap.loss"])
    assert abs(float(_run(fx, O.caption_forward, caption=S["captions"])) - float(g["cap.loss_noprefix"])) < 2e-5 * float(g["cap.loss_noprefix"])


def test_caption_rank_and_generate(fx):
    g = fx[-1]
    r = _run(fx, O.caption_forward, answer=S["classes"], train=False, prefix=S["prefix"], inference="rank", k_test=S["k_test"])
    assert np.array_equal(r.numpy(), g["cap.rank"])
    assert _run(fx, O.caption_forward, train=False, prefix=S["prefix"]) == g["cap.generate"].tolist()


def test_vqa_train_rank_generate(fx):
    g = fx[-1]
    loss = _run(fx, O.vqa_forward, S["questions"], S["answers"], weights=torch.tensor(S["weights"]))
    assert abs(float(loss) - float(g["vqa.loss"])) < 2e-5 * float(g["vqa.loss"])
    r = _run(fx, O.vqa_forward, S["questions"], S["candidates"], train=False, inference="rank", k_test=S["k_test"])
    assert np.array_equal(r.numpy(), g["vqa.rank"])
    assert _run(fx, O.vqa_forward, S["questions"], train=False, inference="generate") == g["vqa.generate"].tolist()


def test_host_tokenisation_helpers_reproduce_the_reference_losses(fx):
    """``prismer_b200.text`` (what the product's ``forward`` and an N3-style data loader call) builds the same ids / masks / labels as
    the reference's ``forward``: scoring them with the oracle decoder gives the reference's own losses."""
    from prismer_b200 import text
    cfg, sd, ex, tok, g = fx
    esd, dsd = O.split_state_dict(sd)
    random.seed(cfg["py_seed"])
    with torch.no_grad():
        enc = O.encoder_forward(ex, esd, cfg["patch"]).transpose(0, 1)
        ids, mask, labels, plen = text.caption_inputs(tok, S["captions"], S["prefix"])
        assert plen == 4 and bool((labels[:, :4] == -100).all()) and ids.shape == mask.shape == labels.shape
        _, loss = O.decoder_forward(ids, mask, enc, dsd, HEADS, labels)
        assert abs(float(loss.mean()) - float(g["cap.loss"])) < 2e-5 * float(g["cap.loss"])
        ids, mask, labels, plen = text.caption_inputs(tok, S["captions"])
        assert plen == 0
        _, loss = O.decoder_forward(ids, mask, enc, dsd, HEADS, labels)
        assert abs(float(loss.mean()) - float(g["cap.loss_noprefix"])) < 2e-5 * float(g["cap.loss_noprefix"])
        ids, mask, labels = text.vqa_inputs(tok, S["questions"], S["answers"])
        _, loss = O.decoder_forward(ids, mask, enc, dsd, HEADS, labels)
        assert abs(float((torch.tensor(S["weights"]) * loss).mean()) - float(g["vqa.loss"])) < 2e-5 * float(g["vqa.loss"])
    q = text.vqa_question(tok, S["questions"])
    assert q.input_ids[0, 0] == 0 and int((q.input_ids == 2).sum()) == 0           # <s> prepended, no </s> (prismer_vqa.py:18-20)
