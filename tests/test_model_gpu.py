"""-m gpu: the assembled CUDA path (through the C-ABI) against (1) the committed golden vectors produced by the
UNMODIFIED reference and (2) the CPU oracle on the same seeded weights / inputs.

Tolerances (stated, not aspirational): the product computes in bf16 storage with fp32 accumulation / statistics, the
goldens and the oracle are fp32.  Each bf16 rounding contributes 2^-9 relative error; through ~6 (tiny) to ~30 (BASE)
layers the measured relative L2 error of encoder states / logits is 0.3-1.0e-2, of gradients 1-3e-2.  Bounds below are
~2x the values measured on B200 and are printed by the tests.  Integer outputs (token ids) must match exactly."""
import random

import numpy as np
import pytest
import torch

from prismer_b200 import synthetic
from tests.helpers import TINY_DEC, build_model, load_golden, rel_l2

pytestmark = pytest.mark.gpu
FULL = synthetic.DEFAULT_EXPERTS


def _cuda_experts(ex):
    return synthetic.experts_to(ex, "cuda")


def test_tiny_eval_matches_reference_golden_A():
    from prismer_b200 import engine
    cfg, g = load_golden("A")
    m, _ = build_model(cfg["width"], cfg["layers"], cfg["patch"], cfg["res"], FULL, TINY_DEC, cfg["seed"])
    m.eval()
    ex = _cuda_experts(synthetic.synth_experts(cfg["B"], cfg["res"], FULL, cfg["label"], cfg["in_seed"]))
    ids, mask = synthetic.synth_tokens(cfg["B"], cfg["T"], TINY_DEC["vocab_size"], cfg["in_seed"], ragged=True)
    random.seed(cfg["py_seed"])
    enc = m.expert_encoder(ex)                                  # [S,B,D]
    e_err = rel_l2(enc.float().cpu(), g["enc"])
    out = m.text_decoder(ids.cuda(), attention_mask=mask.cuda(), encoder_hidden_states=enc.transpose(0, 1),
                         labels=torch.from_numpy(g["labels"]).cuda())
    l_err = rel_l2(out.logits.cpu(), g["logits"])
    loss_err = rel_l2(out.loss.cpu(), g["loss"])
    print(f"tiny-A eval: enc {e_err:.2e} logits {l_err:.2e} loss {loss_err:.2e}")
    assert e_err < 1.5e-2 and l_err < 2e-2 and loss_err < 5e-3
    gen = m.text_decoder.generate(input_ids=torch.from_numpy(g["prefix"]).cuda(), encoder_hidden_states=enc.transpose(0, 1),
                                  num_beams=1, max_length=12, min_length=8)
    assert np.array_equal(gen.cpu().numpy(), g["greedy"]), (gen.cpu().numpy(), g["greedy"])


def test_tiny_train_loss_and_grads_match_reference_golden_A():
    from prismer_b200 import engine
    cfg, _ = load_golden("A")
    _, g = load_golden("A_train")
    m, _ = build_model(cfg["width"], cfg["layers"], cfg["patch"], cfg["res"], FULL, TINY_DEC, cfg["seed"])
    m.expert_encoder.train(); m.text_decoder.eval()             # BN batch statistics on, dropout off (as the fixture)
    ex = _cuda_experts(synthetic.synth_experts(cfg["B"], cfg["res"], FULL, cfg["label"], cfg["in_seed"]))
    ids, mask = synthetic.synth_tokens(cfg["B"], cfg["T"], TINY_DEC["vocab_size"], cfg["in_seed"], ragged=True)
    labels = ids.masked_fill(ids == 1, -100); labels[:, :3] = -100
    random.seed(cfg["py_seed"])
    loss = engine.train_loss(m, ex, ids.cuda(), mask.cuda(), labels.cuda())
    loss.backward()
    torch.cuda.synchronize()
    print(f"tiny-A train: loss {float(loss):.5f} vs {float(g['loss']):.5f}")
    assert abs(float(loss) - float(g["loss"])) / float(g["loss"]) < 5e-3
    named = dict(m.named_parameters())
    worst, bad = 0.0, []
    for k, ref in g.items():
        if not k.startswith("g."):
            continue
        name = ("expert_encoder." + k[4:]) if k[2] == "E" else ("text_decoder." + k[4:])
        grad = named[name].grad.float().cpu()
        nerr = abs(float(grad.norm()) - ref[0]) / max(ref[0], 1e-12)
        verr = rel_l2(grad.flatten()[:2048], ref[1:])
        worst = max(worst, verr)
        print(f"  grad {name}: |g| rel {nerr:.2e}  first-2048 rel-L2 {verr:.2e}")
        # inside the conv stems a handful of ReLU-mask flips (bf16 vs fp32 pre-activations) dominate noise-like gradient
        # sums (see tests/test_stem_gpu.py, which checks the same kernels against a reference on the bf16 grid)
        lim = 0.25 if ("conv1." in name and "rgb" not in name and ".13." not in name) else 6e-2
        bad.append(name) if not (nerr < 5e-2 and verr < lim) else None
    assert not bad, bad
    # BatchNorm running statistics were updated like nn.BatchNorm2d(momentum=0.1)
    sd = m.expert_encoder.state_dict()
    for k, ref in g.items():
        if k.startswith("bn."):
            assert rel_l2(sd[k[3:]].float().cpu(), ref) < 5e-3, k


@pytest.mark.parametrize("name,experts", [("B", ["depth", "seg_coco", "obj_detection"]), ("C", ["normal", "edge", "ocr_detection"])])
def test_tiny_encoder_resample_and_posinterp_paths(name, experts):
    cfg, g = load_golden(name)
    m, _ = build_model(cfg["width"], cfg["layers"], cfg["patch"], cfg["res"], experts, None, cfg["seed"])
    m.eval()
    ex = _cuda_experts(synthetic.synth_experts(cfg["B"], cfg["res"], experts, cfg["label"], cfg["in_seed"]))
    random.seed(cfg["py_seed"])
    enc = m.expert_encoder(ex)
    err = rel_l2(enc.float().cpu(), g["enc"])
    print(f"tiny-{name}: enc {err:.2e}")
    assert err < 1.5e-2


def test_prismerz_greedy_ids_match_reference_golden_Z():
    cfg, g = load_golden("Z")
    m, _ = build_model(cfg["width"], cfg["layers"], cfg["patch"], cfg["res"], [], TINY_DEC, cfg["seed"])
    m.eval()
    ex = _cuda_experts(synthetic.synth_experts(1, cfg["res"], [], 64, cfg["in_seed"]))
    enc = m.expert_encoder(ex)
    assert rel_l2(enc.float().cpu(), g["enc"]) < 1.5e-2
    gen = m.text_decoder.generate(input_ids=torch.from_numpy(g["prefix"]).cuda(), encoder_hidden_states=enc.transpose(0, 1),
                                  num_beams=1, max_length=20, min_length=8)
    assert np.array_equal(gen.cpu().numpy(), g["greedy"])
