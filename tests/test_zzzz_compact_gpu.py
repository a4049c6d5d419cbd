"""-m gpu: compact expert inputs (csrc/compact_inputs.cu) -- uint8 label map + table expanded on the GPU
instead of the reference's CPU in-painting (dataset/utils.py:117-160).

  * ``prismer_expand_labels`` reproduces the reference's ``post_label_process`` outputs (tests/golden/prismer_labels.npz) bit
    for bit -- a pure gather;
  * ``prismer_label_resample`` (in-painting fused with UpsamplingBilinear2d) equals ``prismer_resample_bilinear`` applied to the
    expanded tensor (same fp32 expression on the same values; at most one bf16 ulp apart should the compiler contract the two
    kernels differently, reported);
  * the encoder fed compact inputs returns the same states as the encoder fed the expanded float maps."""
import random

import numpy as np
import pytest
import torch

from prismer_b200 import data, synthetic
from tests.helpers import GOLD, build_model, label_case

pytestmark = pytest.mark.gpu


def test_expand_labels_matches_reference_post_label_process():
    z = dict(np.load(f"{GOLD}/prismer_labels.npz"))
    feats = {k[5:]: torch.from_numpy(v) for k, v in z.items() if k.startswith("feat.")}
    for case in range(3):
        u8, info = label_case(case)
        res = data.compact_label_process(dict(u8), info, feats)
        for k in u8:
            v = res[k]
            cm = (v["label"] if isinstance(v, dict) else v).to("cuda")
            want = z[f"c{case}.{k}.label"] if isinstance(v, dict) else z[f"c{case}.{k}"]
            assert np.array_equal(cm.expand().cpu().numpy(), want), (case, k)          # single sample, shared table
    batch = data.collate_experts([data.compact_label_process(dict(label_case(c)[0]), label_case(c)[1], feats) for c in range(3)])
    for k in ("seg_coco", "normal", "ocr_detection"):                                   # batched, per-image tables
        got = batch[k].to("cuda").expand().cpu()
        assert torch.equal(got, batch[k].expand_on_host()), k


@pytest.mark.parametrize("Hi,Ho", [(224, 56), (224, 64), (64, 16), (24, 6)])
def test_label_resample_equals_resample_of_expansion(Hi, Ho):
    from prismer_b200 import ops
    rs = np.random.RandomState(Hi + Ho)
    B = 3
    u8 = torch.from_numpy(np.where(rs.uniform(size=(B, 1, Hi, Hi)) < 0.2, 255, synthetic._blocky(rs, B, Hi, 9, 40)[:, None]).astype(np.uint8)).cuda()
    for table in (torch.from_numpy(rs.standard_normal((256, 64)).astype(np.float32)).cuda(),
                  torch.from_numpy(rs.standard_normal((B, 256, 64)).astype(np.float32)).cuda()):
        want = ops.resample_bilinear(ops.expand_labels(u8, table), Ho, Ho)
        got = ops.label_resample(u8, table, Ho, Ho)
        torch.cuda.synchronize()
        diff = (got.float() - want.float()).abs()
        exact = float((got.view(torch.int16) == want.view(torch.int16)).float().mean())
        print(f"label_resample {Hi}->{Ho}: {exact * 100:.3f}% of outputs bit-identical, max |diff| {float(diff.max()):.3e}")
        assert float((diff / want.float().abs().clamp_min(1e-3)).max()) <= 2 ** -7         # <= 1 bf16 ulp
        assert exact > 0.999


def test_encoder_on_compact_inputs_equals_encoder_on_float_maps():
    experts = synthetic.DEFAULT_EXPERTS
    m, _ = build_model(256, 2, 16, 64, experts, None, 7)
    m.eval()
    cex = synthetic.synth_compact_experts(2, 64, experts, 64, seed=3)
    fex = synthetic.expand_compact_on_host(cex)
    random.seed(5)
    a = m.expert_encoder(synthetic.experts_to(cex, "cuda")).float()
    random.seed(5)
    b = m.expert_encoder(synthetic.experts_to(fex, "cuda")).float()
    torch.cuda.synchronize()
    err = float((a - b).norm() / b.norm())
    print(f"encoder(compact) vs encoder(float maps): rel-L2 {err:.2e}, bit-identical {bool(torch.equal(a, b))}")
    assert err < 1e-3
