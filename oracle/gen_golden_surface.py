"""Golden values from the reference's own model-level glue: the UNMODIFIED ``PrismerCaption.forward``
(model/prismer_caption.py:15-112) and ``PrismerVQA.forward`` (model/prismer_vqa.py:16-113) -- string tokenisation, label /
prompt masking, answer weighting, first-token top-k + tiled second pass of ``inference='rank'``, beam-3 ``generate`` and
the string decoding -- run on the tiny fixture-A modules.

``Prismer.__init__`` (model/prismer.py:16-37) cannot run here (it downloads the tokenizer vocabulary, CLIP and RoBERTa
weights), so the instances are created without it and handed the three attributes ``forward`` uses: ``expert_encoder`` /
``text_decoder`` (reference modules, seeded weights) and ``tokenizer`` (the deterministic ``HashTokenizer`` stand-in, the
same one the product falls back to without a vocabulary).  Everything executed after that is the reference's code.

TEST INFRASTRUCTURE: needs /root/reference; run here, never on the GPU box.

    python oracle/gen_golden_surface.py      ->  tests/golden/prismer_tiny_surface.npz
"""
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import reference_shim  # noqa: E402
from oracle.gen_golden import GOLD, TINY_DEC, build  # noqa: E402
from prismer_b200 import synthetic  # noqa: E402
from prismer_b200.tokenizer import HashTokenizer  # noqa: E402
from tests.helpers import SURFACE  # noqa: E402


def bare(cls, vit, dec):
    m = cls.__new__(cls)
    torch.nn.Module.__init__(m)
    m.expert_encoder, m.text_decoder, m.tokenizer = vit, dec, HashTokenizer(TINY_DEC["vocab_size"])
    return m.eval()


def main():
    ns = reference_shim.load()
    import importlib
    cap_cls = importlib.import_module("model.prismer_caption").PrismerCaption
    vqa_cls = importlib.import_module("model.prismer_vqa").PrismerVQA
    S = SURFACE
    cfg = S["cfg"]
    full = synthetic.DEFAULT_EXPERTS
    vit, dec, _ = build(ns, cfg["width"], cfg["layers"], cfg["patch"], cfg["res"], full, cfg["seed"])
    ex = synthetic.synth_experts(cfg["B"], cfg["res"], full, cfg["label"], cfg["in_seed"])
    cap, vqa = bare(cap_cls, vit, dec), bare(vqa_cls, vit, dec)
    out = {}
    seed = lambda: random.seed(cfg["py_seed"])          # instance-embedding draw of the encoder (vit.py:141-148)
    with torch.no_grad():
        seed(); out["cap.loss"] = cap(ex, S["captions"], prefix=S["prefix"]).numpy()
        seed(); out["cap.loss_noprefix"] = cap(ex, S["captions"]).numpy()
        seed(); out["cap.rank"] = cap(ex, answer=S["classes"], train=False, prefix=S["prefix"], inference="rank", k_test=S["k_test"]).numpy()
        seed(); out["cap.generate"] = np.array(cap(ex, train=False, prefix=S["prefix"]))
        seed(); out["vqa.loss"] = vqa(ex, S["questions"], S["answers"], weights=torch.tensor(S["weights"])).numpy()
        seed(); out["vqa.rank"] = vqa(ex, S["questions"], S["candidates"], train=False, inference="rank", k_test=S["k_test"]).numpy()
        seed(); out["vqa.generate"] = np.array(vqa(ex, S["questions"], train=False, inference="generate"))
    for k, v in out.items():
        print(k, v)
    np.savez_compressed(os.path.join(GOLD, "prismer_tiny_surface.npz"), **out)


if __name__ == "__main__":
    main()
