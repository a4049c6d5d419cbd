"""CPU: host-side logic of the reference-facing surface (no CUDA needed): freeze policy, expert channel table, tokenizer
stand-in, label masking, deterministic synthetic data, positional-embedding interpolation matrix, greedy output trimming."""
import hashlib

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import reference_shim
from prismer_b200 import synthetic
from tests.helpers import TINY_DEC

TINY = {"roberta_model": dict(TINY_DEC, model_name="roberta-tiny"), "vit_model": "tiny", "vit_dims": [16, 256, 2]}
EXPERTS = ["depth", "normal", "seg_coco", "edge", "obj_detection", "ocr_detection"]


def _model(freeze):
    from prismer_b200.prismer_caption import PrismerCaption
    return PrismerCaption({"experts": EXPERTS, "prismer_model": "tiny", "image_resolution": 64, "freeze": freeze, "prismer_config": TINY})


def test_expert_channels_and_prismerz_quirk():
    from prismer_b200.prismer import expert_channels
    assert expert_channels(EXPERTS) == {"rgb": 3, "depth": 1, "normal": 3, "seg": 64, "edge": 1, "obj_detection": 64, "ocr_detection": 64}
    assert expert_channels("none") == {"rgb": 3}      # iterating the string matches nothing (model/prismer.py:19-27) -> PrismerZ


@pytest.mark.parametrize("mode", ["none", "freeze_lang", "freeze_vision", "freeze_lang_vision"])
def test_freeze_policy_matches_reference_rule(mode):
    """model/prismer.py:39-59, restated literally here as the checker."""
    m = _model(mode)
    for name, p in m.named_parameters():
        lang = "encoder.layer" in name and all(k not in name for k in ["1.self", "1.output", "adaptor"])
        vis = "transformer.resblocks" in name and "adaptor" not in name
        frozen = {"none": False, "freeze_lang": lang, "freeze_vision": vis, "freeze_lang_vision": lang or vis}[mode]
        assert p.requires_grad == (not frozen), (mode, name)
    if mode == "none":
        assert m.ignored_modules is None
    else:
        assert len(m.ignored_modules) > 0


def test_trainable_count_of_base_matches_survey():
    from prismer_b200.prismer_caption import PrismerCaption
    with torch.device("meta"):
        m = PrismerCaption({"experts": EXPERTS, "prismer_model": "prismer_base", "image_resolution": 224, "freeze": "freeze_vision"})
    total = sum(p.numel() for p in m.parameters())
    train = sum(p.numel() for p in m.parameters() if p.requires_grad)
    assert round(total / 1e6, 1) == 327.5 and round(train / 1e6, 1) == 242.4       # SURVEY.md section 8 / BASELINE.md section 3


def test_hash_tokenizer_surface_and_label_masking():
    from prismer_b200.tokenizer import HashTokenizer
    tok = HashTokenizer(50265)
    a = tok("A picture of")
    assert a.input_ids[0] == 0 and a.input_ids[-1] == 2 and len(a.input_ids) == 5
    b = tok(["A picture of a dog", "hi"], padding="longest", return_tensors="pt")
    assert b.input_ids.shape == (2, 7) and b.input_ids[1, 3:].tolist() == [1, 1, 1, 1] and b.attention_mask[1].tolist() == [1, 1, 1, 0, 0, 0, 0]
    assert tok("A picture of a dog").input_ids[:4] == a.input_ids[:4]          # prefix tokens are a prefix of the caption's
    labels = b.input_ids.masked_fill(b.input_ids == tok.pad_token_id, -100)    # prismer_caption.py:22-26
    labels[:, :len(a.input_ids) - 1] = -100
    assert (labels[1] == -100).all() and labels[0, 4:].tolist() == b.input_ids[0, 4:].tolist()


def test_synthetic_data_is_bit_stable():
    """The fixtures depend on these streams: numpy RandomState keyed by crc32(name)."""
    t = synthetic.synth_tensor("expert_encoder.ln_pre.weight", (8,), 7)
    assert hashlib.sha1(t.numpy().tobytes()).hexdigest() == hashlib.sha1(synthetic.synth_tensor("expert_encoder.ln_pre.weight", (8,), 7).numpy().tobytes()).hexdigest()
    ex = synthetic.synth_experts(1, 32, ["depth", "obj_detection"], 32, 3)
    assert list(ex) == ["rgb", "depth", "obj_detection"] and ex["obj_detection"]["instance"].dtype == torch.int64
    assert set(np.unique(ex["obj_detection"]["instance"].numpy())) <= {0, 1, 2, 3, 255}
    lab = ex["obj_detection"]["label"]
    assert lab.shape == (1, 64, 32, 32) and torch.equal(lab[0, :, 0, 0], lab[0, :, 1, 1])      # piece-wise constant features
    ids, mask = synthetic.synth_tokens(3, 10, 1000, 5, ragged=True)
    assert (ids[:, 0] == 0).all() and ((ids == 1) == (mask == 0)).all()


@pytest.mark.parametrize("P,n", [(900, 196), (1156, 256), (49, 16)])
def test_interpolation_matrix_equals_bicubic_interpolate(P, n):
    """utils.py:34-44 is linear in the embedding: the fixed matrix reproduces F.interpolate(bicubic, align_corners=False)."""
    from prismer_b200.modules.utils import interpolate_pos_embed, interpolation_matrix
    pos = torch.randn(P, 24)
    o, m = int(P ** 0.5), int(n ** 0.5)
    ref = F.interpolate(pos.reshape(1, o, o, -1).permute(0, 3, 1, 2), size=(m, m), mode="bicubic", align_corners=False)
    ref = ref.permute(0, 2, 3, 1).flatten(0, 2)
    assert torch.allclose(interpolation_matrix(P, n) @ pos, ref, atol=1e-5)
    assert torch.allclose(interpolate_pos_embed(pos, n), ref, atol=1e-5)
    assert interpolate_pos_embed(pos, P) is pos


def test_trim_finished_matches_hf_stopping_rule():
    from prismer_b200.generation import trim_finished
    ids = torch.tensor([[0, 5, 6, 7, 9, 2, 1, 1], [0, 5, 6, 7, 8, 8, 2, 1]])
    assert trim_finished(ids, 4, 2).shape[1] == 7       # stop right after the slowest row emitted </s>
    assert trim_finished(torch.tensor([[0, 5, 6, 7, 9, 9, 9, 9]]), 4, 2).shape[1] == 8


@pytest.mark.skipif(not reference_shim.available(), reason="reference tree not mounted")
def test_state_dict_keys_equal_live_reference():
    from prismer_b200 import modeling
    from prismer_b200.prismer import expert_channels
    ns = reference_shim.load()
    ref_v = ns.VisionTransformer(64, 16, 256, 2, 4, expert_channels(EXPERTS))
    ref_d = ns.build_decoder(TINY_DEC)
    mine_v, mine_d = modeling.build_encoder(256, 2, 16, 64, EXPERTS), modeling.build_decoder(TINY_DEC)
    for a, b in ((ref_v, mine_v), (ref_d, mine_d)):
        sa = {k: tuple(v.shape) for k, v in a.state_dict().items()}
        sb = {k: tuple(v.shape) for k, v in b.state_dict().items()}
        assert sa == sb
        assert {k for k, _ in a.named_parameters()} == {k for k, _ in b.named_parameters()}
