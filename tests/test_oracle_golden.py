"""Pins ``oracle/prismer_oracle.py`` (the CPU restatement) against the committed golden vectors that
``oracle/gen_golden.py`` produced from the UNMODIFIED reference modules."""
import random

import numpy as np
import pytest
import torch

from oracle import prismer_oracle as O
from prismer_b200 import synthetic
from prismer_b200.modeling import template_state_dict
from tests.helpers import TINY_DEC, load_golden, rel_l2

FULL = synthetic.DEFAULT_EXPERTS
TOL = 2e-5  # fp32 CPU restatement vs fp32 CPU reference: summation-order noise only


def _sd(cfg, experts, dec=True):
    tmpl = template_state_dict(width=cfg["width"], layers=cfg["layers"], patch=cfg["patch"], res=cfg["res"],
                               experts=experts, dec_cfg=TINY_DEC if dec else None)
    return synthetic.synth_state_dict(tmpl, cfg["seed"])


def test_encoder_decoder_eval_A():
    cfg, g = load_golden("A")
    sd = _sd(cfg, FULL)
    ex = synthetic.synth_experts(cfg["B"], cfg["res"], FULL, cfg["label"], cfg["in_seed"])
    ids, mask = synthetic.synth_tokens(cfg["B"], cfg["T"], TINY_DEC["vocab_size"], cfg["in_seed"], ragged=True)
    assert np.array_equal(ids.numpy(), g["ids"]) and np.array_equal(mask.numpy(), g["mask"])
    esd, dsd = O.split_state_dict(sd)
    random.seed(cfg["py_seed"])
    with torch.no_grad():
        enc = O.encoder_forward(ex, esd, cfg["patch"])
        assert rel_l2(enc, g["enc"]) < TOL
        logits, loss = O.decoder_forward(ids, mask, enc.transpose(0, 1), dsd, TINY_DEC["num_attention_heads"],
                                         torch.from_numpy(g["labels"]))
        assert rel_l2(logits, g["logits"]) < TOL
        assert rel_l2(loss, g["loss"]) < TOL
        out, _ = O.greedy_generate(enc.transpose(0, 1), torch.from_numpy(g["prefix"]), dsd,
                                   TINY_DEC["num_attention_heads"], max_length=12, min_length=8)
    assert np.array_equal(out.numpy(), g["greedy"])  # token ids: bit-exact


def test_train_mode_bn_and_grads_A():
    cfg, _ = load_golden("A")
    _, g = load_golden("A_train")
    sd = _sd(cfg, FULL)
    for v in sd.values():
        if v.dtype.is_floating_point:
            v.requires_grad_(True)
    ex = synthetic.synth_experts(cfg["B"], cfg["res"], FULL, cfg["label"], cfg["in_seed"])
    ids, mask = synthetic.synth_tokens(cfg["B"], cfg["T"], TINY_DEC["vocab_size"], cfg["in_seed"], ragged=True)
    random.seed(cfg["py_seed"])
    loss, _, enc = O.caption_train_loss(ex, ids, mask, 3, sd, cfg["patch"], TINY_DEC["num_attention_heads"], training_bn=True)
    assert rel_l2(enc.transpose(0, 1), g["enc"]) < TOL
    assert rel_l2(loss, g["loss"]) < TOL
    loss.backward()
    for k, ref in g.items():
        if not k.startswith("g."):
            continue
        name = k[2:].replace("E.", "expert_encoder.", 1) if k[2] == "E" else k[2:].replace("D.", "text_decoder.", 1)
        grad = sd[name].grad
        if name.endswith("word_embeddings.weight"):
            pass  # tied: the oracle's single tensor receives embedding + LM-head gradient, like the reference
        assert abs(float(grad.norm()) - ref[0]) <= 1e-4 * max(ref[0], 1e-6), name
        assert rel_l2(grad.flatten()[:2048], ref[1:]) < 2e-4, name


@pytest.mark.parametrize("name,experts", [("B", ["depth", "seg_coco", "obj_detection"]),
                                          ("C", ["normal", "edge", "ocr_detection"])])
def test_encoder_resample_paths(name, experts):
    cfg, g = load_golden(name)
    sd = _sd(cfg, experts, dec=False)
    ex = synthetic.synth_experts(cfg["B"], cfg["res"], experts, cfg["label"], cfg["in_seed"])
    esd, _ = O.split_state_dict(sd)
    random.seed(cfg["py_seed"])
    with torch.no_grad():
        enc = O.encoder_forward(ex, esd, cfg["patch"])
    assert rel_l2(enc, g["enc"]) < TOL


def test_prismerz_greedy_Z():
    cfg, g = load_golden("Z")
    sd = _sd(cfg, [])
    ex = synthetic.synth_experts(1, cfg["res"], [], 64, cfg["in_seed"])
    esd, dsd = O.split_state_dict(sd)
    with torch.no_grad():
        enc = O.encoder_forward(ex, esd, cfg["patch"])
        assert rel_l2(enc, g["enc"]) < TOL
        out, _ = O.greedy_generate(enc.transpose(0, 1), torch.from_numpy(g["prefix"]), dsd, TINY_DEC["num_attention_heads"])
    assert np.array_equal(out.numpy(), g["greedy"])
