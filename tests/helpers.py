"""Shared helpers for the test-suite (test infrastructure; may import ``oracle``)."""
import os

import numpy as np
import torch

from prismer_b200 import synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

TINY_DEC = {
    "attention_probs_dropout_prob": 0.1, "bos_token_id": 0, "eos_token_id": 2, "hidden_act": "gelu",
    "hidden_dropout_prob": 0.1, "hidden_size": 256, "vision_hidden_size": 256, "initializer_range": 0.02,
    "intermediate_size": 1024, "layer_norm_eps": 1e-05, "max_position_embeddings": 514,
    "model_name": "roberta-tiny", "num_attention_heads": 4, "num_hidden_layers": 2, "pad_token_id": 1,
    "type_vocab_size": 1, "vocab_size": 1000, "num_decoder_layers": 4, "is_decoder": True,
}


def load_golden(name):
    z = np.load(os.path.join(GOLD, f"prismer_tiny_{name}.npz"))
    cfg = {k[4:]: z[k].item() for k in z.files if k.startswith("cfg.")}
    return cfg, {k: z[k] for k in z.files if not k.startswith("cfg.")}


def rel_l2(a, b):
    a = torch.as_tensor(a).double().flatten()
    b = torch.as_tensor(b).double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


class Holder(torch.nn.Module):
    """expert_encoder + text_decoder pair with the reference's attribute names (a config-free ``Prismer``)."""

    def __init__(self, enc, dec=None):
        super().__init__()
        self.expert_encoder = enc
        if dec is not None:
            self.text_decoder = dec


def build_model(width, layers, patch, res, experts, dec_cfg, seed, device="cuda", freeze=None):
    from prismer_b200 import modeling
    enc = modeling.build_encoder(width, layers, patch, res, experts)
    dec = modeling.build_decoder(dec_cfg) if dec_cfg is not None else None
    m = Holder(enc, dec)
    sd = synthetic.synth_state_dict(m.state_dict(), seed)
    m.load_state_dict(sd)
    if freeze == "freeze_vision":
        for n, p in m.named_parameters():
            p.requires_grad = not ("transformer.resblocks" in n and "adaptor" not in n)
    m.to(device)
    return m, sd


# ------------------------------------------------------------------------------------------------ beam-search fixtures
# Cases of tests/golden/prismer_tiny_beam.npz (made by oracle/gen_golden_beam.py from the reference decoder).  ``boost`` is
# added to the LM-head eos bias so that hypotheses finish at different lengths; lengths are relative to the prompt length T0.
BEAM_SEED = 21
BEAM_CASES = [
    dict(name="cap_a", B=4, T0=4, ragged=False, S=20, boost=1.5, nb=3, max_add=16, min_add=4, lp=1.0),
    dict(name="cap_b", B=4, T0=4, ragged=False, S=20, boost=2.0, nb=3, max_add=16, min_add=4, lp=1.0),
    dict(name="cap_c", B=3, T0=4, ragged=False, S=12, boost=2.5, nb=3, max_add=16, min_add=4, lp=1.0),
    dict(name="cap_d", B=2, T0=4, ragged=False, S=12, boost=2.0, nb=4, max_add=12, min_add=0, lp=2.0),
    dict(name="cap_e", B=3, T0=4, ragged=False, S=12, boost=0.0, nb=3, max_add=8, min_add=4, lp=1.0),      # nothing finishes early
    dict(name="vqa_a", B=4, T0=7, ragged=True, S=20, boost=2.0, nb=3, max_add=10, min_add=2, lp=-1.0),
    dict(name="vqa_b", B=3, T0=6, ragged=True, S=12, boost=1.0, nb=3, max_add=10, min_add=2, lp=-1.0),
]


# greedy decoding of right-padded (VQA-style) and uniform (caption-style) prompts: oracle/gen_golden_greedy.py
GREEDY_CASES = [
    dict(name="g_vqa_a", B=4, T0=7, ragged=True, S=20, boost=2.0, max_add=10, min_add=2),
    dict(name="g_vqa_b", B=3, T0=6, ragged=True, S=12, boost=3.5, max_add=10, min_add=2),
    dict(name="g_vqa_c", B=4, T0=8, ragged=True, S=12, boost=0.0, max_add=6, min_add=6),
    dict(name="g_cap_a", B=4, T0=4, ragged=False, S=20, boost=3.0, max_add=16, min_add=4),
]


def load_greedy_golden():
    z = np.load(os.path.join(GOLD, "prismer_tiny_greedy.npz"))
    return {k: z[k] for k in z.files}


def beam_decoder_state(template, boost):
    """Decoder state_dict (keys without the ``text_decoder.`` prefix) for a beam case: seeded fill + eos-bias boost."""
    sd = synthetic.synth_state_dict({"text_decoder." + k: v for k, v in template.items()}, BEAM_SEED)
    sd = {k[len("text_decoder."):]: v for k, v in sd.items()}
    sd["lm_head.bias"] = sd["lm_head.bias"].clone()
    sd["lm_head.bias"][TINY_DEC["eos_token_id"]] += boost
    sd["lm_head.decoder.bias"] = sd["lm_head.bias"]
    return sd


def beam_case_inputs(c):
    """(prompt ids, prompt mask, visual tokens [B, S, Dv]) of a beam case."""
    ids, mask = synthetic.synth_tokens(c["B"], c["T0"] + 1, TINY_DEC["vocab_size"], BEAM_SEED + len(c["name"]) + c["B"],
                                       ragged=c["ragged"])
    if c["ragged"]:                                       # VQA: `<s> question </s>` right-padded (prismer_vqa.py:19,46-47)
        ids, mask = ids[:, :c["T0"]].clone(), mask[:, :c["T0"]].clone()
        ids[:, 0] = 0
    else:                                                 # caption: one prefix for every row, `</s>` dropped (prismer_caption.py:38-40)
        ids = ids[:1, :c["T0"]].repeat(c["B"], 1)
        ids[ids == 2] = 7
        mask = torch.ones_like(ids)
    rs = np.random.RandomState(BEAM_SEED * 1000 + sum(map(ord, c["name"])))      # O(1) visual tokens, like ln_post output
    enc = torch.from_numpy(rs.standard_normal((c["B"], c["S"], TINY_DEC["vision_hidden_size"])).astype(np.float32))
    return ids, mask, enc


def load_beam_golden():
    z = np.load(os.path.join(GOLD, "prismer_tiny_beam.npz"))
    return {k: z[k] for k in z.files}


def decoder_template():
    from prismer_b200.modeling import template_state_dict
    t = template_state_dict(width=256, layers=2, patch=16, res=64, experts=[], dec_cfg=TINY_DEC)
    return {k[len("text_decoder."):]: v for k, v in t.items() if k.startswith("text_decoder.")}


# ------------------------------------------------------------------------------------------------ model-level (string API) fixture
# Inputs of tests/golden/prismer_tiny_surface.npz (oracle/gen_golden_surface.py: the reference's own PrismerCaption / PrismerVQA
# ``forward`` on the fixture-A modules with the HashTokenizer stand-in).
SURFACE = dict(
    cfg=dict(width=256, layers=2, patch=16, res=64, label=64, B=2, seed=7, in_seed=11, py_seed=1234),
    prefix="A picture of",
    captions=["A picture of a dog sleeping on a red couch", "A picture of two people riding bikes"],
    classes=["dog", "cat", "traffic light", "people", "a couch in a room"],
    questions=["what is the animal doing", "how many people are there in the picture"],
    answers=["sleeping", "two people"],
    weights=[0.5, 1.0],
    candidates=["sleeping", "running fast", "two people", "no", "yes it is", "a dog"],
    k_test=3,
)


# ------------------------------------------------------------------------------------------------ expert-label fixtures
def label_case(case: int, size: int = 24):
    """Seeded uint8 expert maps of one sample + the ``labels_info`` side data (dataset/utils.py:98-110) for
    tests/golden/prismer_labels.npz.  Case 2 has a constant depth map (max == min) and an all-background ocr map."""
    rs = np.random.RandomState(100 + case)
    u8 = {}
    u8["depth"] = torch.from_numpy(rs.randint(3, 250, (1, size, size)).astype(np.uint8))
    u8["normal"] = torch.from_numpy(rs.randint(0, 256, (3, size, size)).astype(np.uint8))
    u8["edge"] = torch.from_numpy(rs.randint(0, 120, (1, size, size)).astype(np.uint8))
    if case == 2:
        u8["depth"][:] = 77
    def ids(n, cells, bg_frac):
        m = synthetic._blocky(rs, 1, size, cells, n).astype(np.int64)
        m[rs.uniform(size=m.shape) < bg_frac] = 255
        return torch.from_numpy(m.astype(np.uint8))
    u8["seg_coco"] = torch.from_numpy(np.where(rs.uniform(size=(1, size, size)) < 0.2, 255,
                                               synthetic._blocky(rs, 1, size, 5, 133)).astype(np.uint8))
    u8["seg_ade"] = torch.from_numpy(np.where(rs.uniform(size=(1, size, size)) < 0.1, 255,
                                              synthetic._blocky(rs, 1, size, 4, 150)).astype(np.uint8))
    u8["obj_detection"] = ids(5, 4, 0.3)
    u8["ocr_detection"] = ids(3, 3, 0.5)
    if case == 2:
        u8["ocr_detection"][:] = 255
    info = {"obj_detection": {str(i): int(rs.randint(0, 32)) for i in range(5)},
            "ocr_detection": {i: {"features": torch.from_numpy(rs.standard_normal(64).astype(np.float32))} for i in range(3)}}
    return u8, info
