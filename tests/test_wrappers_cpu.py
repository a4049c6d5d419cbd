"""CPU: the product's task wrappers (``PrismerCaption`` / ``PrismerVQA`` ``forward``, ``rank``, ``text.py``, ``generation.py``, the tokenizer
glue) run end to end with the engine's entry points replaced by oracle-backed stand-ins, against values produced by the
reference's OWN ``forward`` methods (tests/golden/prismer_tiny_surface.npz).  Everything above the engine boundary is exercised
exactly as on the GPU; on hardware only the arithmetic behind these five entry points changes (tests/test_zzz_surface_golden_gpu.py)."""
import random
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import prismer_oracle as O
from prismer_b200 import engine, synthetic
from prismer_b200.modules.roberta import CausalLMOutput
from tests.helpers import GOLD, SURFACE as S, TINY_DEC

FULL = synthetic.DEFAULT_EXPERTS
HEADS = TINY_DEC["num_attention_heads"]


@pytest.fixture()
def stand_in(monkeypatch):
    """engine.{train_loss, encoder_apply, decoder_apply, decoder_forward, cross_kv, _store, _experts_check} -> fp32 oracle on the CPU."""
    def sds(root):
        sd = {k: v.detach().float() for k, v in root.state_dict().items()}
        return O.split_state_dict(sd) if any(k.startswith("expert_encoder.") for k in sd) else (None, sd)

    def enc_sd(vit):
        return {k: v.detach().float() for k, v in vit.state_dict().items()}

    def dec_sd(dec):
        return {k: v.detach().float() for k, v in dec.state_dict().items()}

    def train_loss(model, experts, input_ids, attention_mask, labels, weights=None):
        enc = O.encoder_forward(experts, enc_sd(model.expert_encoder), S["cfg"]["patch"]).transpose(0, 1)
        _, loss = O.decoder_forward(input_ids, attention_mask, enc, dec_sd(model.text_decoder), HEADS, labels)
        return (loss if weights is None else weights * loss).mean()

    def encoder_apply(vit, experts):
        return O.encoder_forward(experts, enc_sd(vit), S["cfg"]["patch"])

    def decoder_apply(dec, input_ids, attention_mask, enc, labels=None, weights=None, enc_repeat=1):
        enc = enc.repeat_interleave(enc_repeat, dim=0)          # the reference's tile() (prismer_caption.py:94-96)
        if attention_mask is None:
            attention_mask = torch.ones_like(input_ids)
        logits, loss = O.decoder_forward(input_ids, attention_mask, enc.float(), dec_sd(dec), HEADS, labels)
        return CausalLMOutput(loss=loss, logits=logits)

    def cross_kv(dec, enc):
        return SimpleNamespace(B=enc.shape[0], enc=enc)

    def decoder_forward(dec, input_ids, attention_mask, enc, labels, weights, save, kv=None, last_only=False, **_):
        logits, _ = O.decoder_forward(input_ids, attention_mask, (kv.enc if kv is not None else enc).float(), dec_sd(dec), HEADS)
        return (logits[:, -1].float() if last_only else logits.flatten(0, 1)), None, None, None

    for name, fn in dict(train_loss=train_loss, encoder_apply=encoder_apply, decoder_apply=decoder_apply, cross_kv=cross_kv,
                         decoder_forward=decoder_forward, _experts_check=lambda e: None,
                         _store=lambda m: SimpleNamespace(refresh=lambda: None)).items():
        monkeypatch.setattr(engine, name, fn)
    from prismer_b200 import generation, ops
    monkeypatch.setattr(generation, "KV_CACHE", False)     # the stand-ins replace decoder_forward: the cache-less schedule calls it per step
    monkeypatch.setattr(ops, "argmax", lambda last, V, suppress_eos=False, eos=2: (
        last.masked_fill(torch.arange(last.shape[1]) == eos, -float("inf")) if suppress_eos else last)[:, :V].argmax(-1))


def _model(cls):
    cfg = S["cfg"]
    tiny = {"roberta_model": dict(TINY_DEC, model_name="roberta-tiny"), "vit_model": "tiny", "vit_dims": [cfg["patch"], cfg["width"], cfg["layers"]]}
    m = cls({"experts": FULL, "prismer_model": "tiny", "image_resolution": cfg["res"], "freeze": "none", "prismer_config": tiny})
    m.load_state_dict(synthetic.synth_state_dict(m.state_dict(), cfg["seed"]))
    return m.eval()


def _experts():
    cfg = S["cfg"]
    return synthetic.synth_experts(cfg["B"], cfg["res"], FULL, cfg["label"], cfg["in_seed"])


def test_caption_wrapper_against_reference_forward(stand_in):
    from prismer_b200.prismer_caption import PrismerCaption
    g, m, ex = dict(np.load(f"{GOLD}/prismer_tiny_surface.npz")), _model(PrismerCaption), _experts()
    seed = lambda: random.seed(S["cfg"]["py_seed"])
    with torch.no_grad():
        seed(); assert abs(float(m(ex, S["captions"], prefix=S["prefix"])) - float(g["cap.loss"])) < 2e-5 * float(g["cap.loss"])
        seed(); assert abs(float(m(ex, S["captions"])) - float(g["cap.loss_noprefix"])) < 2e-5 * float(g["cap.loss_noprefix"])
        seed(); r = m(ex, answer=S["classes"], train=False, prefix=S["prefix"], inference="rank", k_test=S["k_test"])
        assert np.array_equal(r.numpy(), g["cap.rank"])
        seed(); assert m(ex, train=False, prefix=S["prefix"]) == g["cap.generate"].tolist()          # beam 3, strings


def test_vqa_wrapper_against_reference_forward(stand_in):
    from prismer_b200 import text
    from prismer_b200.prismer_vqa import PrismerVQA
    g, m, ex = dict(np.load(f"{GOLD}/prismer_tiny_surface.npz")), _model(PrismerVQA), _experts()
    seed = lambda: random.seed(S["cfg"]["py_seed"])
    with torch.no_grad():
        seed(); loss = m(ex, S["questions"], S["answers"], weights=torch.tensor(S["weights"]))
        assert abs(float(loss) - float(g["vqa.loss"])) < 2e-5 * float(g["vqa.loss"])
        ids, mask, labels = text.vqa_inputs(m.tokenizer, S["questions"], S["answers"])               # N3: pre-tokenised tensors
        seed(); loss2 = m(ex, weights=torch.tensor(S["weights"]), input_ids=ids, attention_mask=mask, labels=labels)
        assert float(loss2) == float(loss)
        seed(); r = m(ex, S["questions"], S["candidates"], train=False, inference="rank", k_test=S["k_test"])
        assert np.array_equal(r.numpy(), g["vqa.rank"])
        seed(); assert m(ex, S["questions"], train=False, inference="generate") == g["vqa.generate"].tolist()


def test_greedy_wrapper_path_matches_reference_golden(stand_in):
    """``generate(num_beams=1)`` through generation.greedy on fixture A's greedy golden."""
    from prismer_b200 import modeling
    from tests.helpers import build_model, load_golden
    cfg, g = load_golden("A")
    m, _ = build_model(cfg["width"], cfg["layers"], cfg["patch"], cfg["res"], FULL, TINY_DEC, cfg["seed"], device="cpu")
    m.eval()
    ex = synthetic.synth_experts(cfg["B"], cfg["res"], FULL, cfg["label"], cfg["in_seed"])
    random.seed(cfg["py_seed"])
    with torch.no_grad():
        enc = m.expert_encoder(ex)
        out = m.text_decoder.generate(input_ids=torch.from_numpy(g["prefix"]), encoder_hidden_states=enc.transpose(0, 1), num_beams=1,
                                      max_length=12, min_length=8)
    assert np.array_equal(out.numpy(), g["greedy"])
