"""-m gpu: the product's ``PrismerCaption`` / ``PrismerVQA`` string API against values produced by the reference's OWN
``forward`` methods on the same seeded weights and inputs (tests/golden/prismer_tiny_surface.npz, oracle/gen_golden_surface.py).

Losses: bf16 storage / fp32 accumulation vs the fp32 reference -> relative 5e-3 (measured ~1e-4..1e-3 on the other fixtures).
Rank indices (int64): must equal the reference whenever the fp32 decision margin (oracle log-prob gap between the best two
candidates, and between the k-th and (k+1)-th first-token probability) exceeds the bf16 noise floor."""
import random

import numpy as np
import pytest
import torch

from prismer_b200 import synthetic
from tests.helpers import GOLD, SURFACE as S, TINY_DEC

pytestmark = pytest.mark.gpu
FULL = synthetic.DEFAULT_EXPERTS
HEADS = TINY_DEC["num_attention_heads"]


def _model(cls):
    cfg = S["cfg"]
    tiny = {"roberta_model": dict(TINY_DEC, model_name="roberta-tiny"), "vit_model": "tiny", "vit_dims": [cfg["patch"], cfg["width"], cfg["layers"]]}
    m = cls({"experts": FULL, "prismer_model": "tiny", "image_resolution": cfg["res"], "freeze": "none", "prismer_config": tiny})
    m.load_state_dict(synthetic.synth_state_dict(m.state_dict(), cfg["seed"]))
    return m.cuda().eval()                      # the goldens were taken in eval mode: no dropout, BatchNorm on running stats


def _experts():
    cfg = S["cfg"]
    return synthetic.experts_to(synthetic.synth_experts(cfg["B"], cfg["res"], FULL, cfg["label"], cfg["in_seed"]), "cuda")


def _gold():
    return dict(np.load(f"{GOLD}/prismer_tiny_surface.npz"))


def _oracle_rank(m, start, answers):
    """fp32 oracle rank on the product's weights -> (ids, first-token margin, final margin)."""
    from oracle import prismer_oracle as O
    cfg = S["cfg"]
    sd = {k: v.detach().float().cpu() for k, v in m.state_dict().items()}
    esd, dsd = O.split_state_dict(sd)
    ex = synthetic.synth_experts(cfg["B"], cfg["res"], FULL, cfg["label"], cfg["in_seed"])
    random.seed(cfg["py_seed"])
    with torch.no_grad():
        enc = O.encoder_forward(ex, esd, cfg["patch"]).transpose(0, 1)
        ids, topk, lps = O.rank_answers(enc, start.input_ids, start.attention_mask, answers.input_ids, answers.attention_mask, dsd,
                                        HEADS, S["k_test"])
        logits, _ = O.decoder_forward(start.input_ids, start.attention_mask, enc, dsd, HEADS)
        p = torch.softmax(logits[:, -1], 1).index_select(1, answers.input_ids[:, 0]).sort(dim=1, descending=True).values
    first = (p[:, S["k_test"] - 1] / p[:, S["k_test"]] - 1).min().item() if p.shape[1] > S["k_test"] else 1.0
    top2 = lps.sort(dim=1, descending=True).values
    return ids, first, (top2[:, 0] - top2[:, 1]).min().item()


def test_caption_surface_matches_reference_forward():
    from prismer_b200.prismer_caption import PrismerCaption
    g, m, ex = _gold(), _model(PrismerCaption), _experts()
    seed = lambda: random.seed(S["cfg"]["py_seed"])
    with torch.no_grad():
        seed(); l1 = float(m(ex, S["captions"], prefix=S["prefix"]))
        seed(); l0 = float(m(ex, S["captions"]))
    print(f"caption loss {l1:.4f} vs {float(g['cap.loss']):.4f}; no-prefix {l0:.4f} vs {float(g['cap.loss_noprefix']):.4f}")
    assert abs(l1 - float(g["cap.loss"])) < 5e-3 * float(g["cap.loss"])
    assert abs(l0 - float(g["cap.loss_noprefix"])) < 5e-3 * float(g["cap.loss_noprefix"])
    seed(); r = m(ex, answer=S["classes"], train=False, prefix=S["prefix"], inference="rank", k_test=S["k_test"])
    tok = m.tokenizer
    p = tok([S["prefix"]] * S["cfg"]["B"], padding="longest", return_tensors="pt")
    p.input_ids, p.attention_mask = p.input_ids[:, :-1], p.attention_mask[:, :-1]
    a = tok([" " + x.lower() + "</s>" for x in S["classes"]], padding="longest", return_tensors="pt", add_special_tokens=False)
    ref, first, final = _oracle_rank(m, p, a)
    assert np.array_equal(ref.numpy(), g["cap.rank"])
    print(f"caption rank {r.tolist()} vs reference {g['cap.rank'].tolist()} (margins: first-token {first:.3f}, final {final:.3f})")
    if first > 5e-2 and final > 5e-2:
        assert np.array_equal(r.cpu().numpy(), g["cap.rank"])
    seed(); caps = m(ex, train=False, prefix=S["prefix"])
    assert isinstance(caps, list) and len(caps) == S["cfg"]["B"] and all(isinstance(c, str) and len(c) > 0 for c in caps)
    print("caption generate identical to the reference:", caps == g["cap.generate"].tolist())


def test_vqa_surface_matches_reference_forward():
    from prismer_b200.prismer_vqa import PrismerVQA
    g, m, ex = _gold(), _model(PrismerVQA), _experts()
    seed = lambda: random.seed(S["cfg"]["py_seed"])
    with torch.no_grad():
        seed(); loss = float(m(ex, S["questions"], S["answers"], weights=torch.tensor(S["weights"])))
    print(f"vqa loss {loss:.4f} vs {float(g['vqa.loss']):.4f}")
    assert abs(loss - float(g["vqa.loss"])) < 5e-3 * float(g["vqa.loss"])
    seed(); r = m(ex, S["questions"], S["candidates"], train=False, inference="rank", k_test=S["k_test"])
    tok = m.tokenizer
    q = tok(["<s>" + x.capitalize() for x in S["questions"]], padding="longest", truncation=True, max_length=35, add_special_tokens=False,
            return_tensors="pt")
    a = tok([" " + x.capitalize() + "</s>" for x in S["candidates"]], padding="longest", return_tensors="pt", add_special_tokens=False)
    ref, first, final = _oracle_rank(m, q, a)
    assert np.array_equal(ref.numpy(), g["vqa.rank"])
    print(f"vqa rank {r.tolist()} vs reference {g['vqa.rank'].tolist()} (margins: first-token {first:.3f}, final {final:.3f})")
    if first > 5e-2 and final > 5e-2:
        assert np.array_equal(r.cpu().numpy(), g["vqa.rank"])
    seed(); ans = m(ex, S["questions"], train=False, inference="generate")
    assert isinstance(ans, list) and len(ans) == S["cfg"]["B"] and all(isinstance(x, str) for x in ans)
    print("vqa generate identical to the reference:", ans == g["vqa.generate"].tolist())
