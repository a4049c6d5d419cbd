"""Expert-label post-processing (dataset/utils.py:117-160) without a GPU:

  * the oracle restatement is pinned bit-exactly against the reference's own ``post_label_process`` outputs
    (tests/golden/prismer_labels.npz, oracle/gen_golden_labels.py);
  * the product's compact format (``prismer_b200.data.compact_label_process``: uint8 map + <=256-row table) stands for
    exactly the same tensors: expanding it on the host reproduces the reference output bit for bit, for every expert."""
import numpy as np
import pytest
import torch

from oracle import prismer_oracle as O
from prismer_b200 import data
from tests.helpers import GOLD, label_case


@pytest.fixture(scope="module")
def gold():
    z = dict(np.load(f"{GOLD}/prismer_labels.npz"))
    feats = {k[5:]: torch.from_numpy(v) for k, v in z.items() if k.startswith("feat.")}
    return z, feats


def _float_inputs(u8):
    out = {}
    for k, v in u8.items():
        f = v.to(torch.float32).div(255)                                 # transforms_f.to_tensor (dataset/utils.py:56-63)
        out[k] = f if k in ("depth", "normal", "edge") else (f * 255).long()
    return out


@pytest.mark.parametrize("case", [0, 1, 2])
def test_oracle_post_label_process_matches_reference(case, gold):
    z, feats = gold
    u8, info = label_case(case)
    res = O.post_label_process(_float_inputs(u8), info, feats)
    for k, v in res.items():
        if isinstance(v, dict):
            assert np.array_equal(v["label"].numpy(), z[f"c{case}.{k}.label"]) and np.array_equal(v["instance"].numpy(), z[f"c{case}.{k}.instance"])
        else:
            assert np.array_equal(v.numpy(), z[f"c{case}.{k}"]), k      # bit-exact, floats included


@pytest.mark.parametrize("case", [0, 1, 2])
def test_compact_format_stands_for_the_reference_tensors(case, gold):
    z, feats = gold
    u8, info = label_case(case)
    res = data.compact_label_process(dict(u8), info, feats)
    for k in u8:
        v = res[k]
        cm = v["label"] if isinstance(v, dict) else v
        assert isinstance(cm, data.CompactMap) and cm.u8.dtype == torch.uint8 and cm.table.shape[0] == 256
        want = z[f"c{case}.{k}.label"] if isinstance(v, dict) else z[f"c{case}.{k}"]
        assert tuple(cm.shape) == want.shape
        assert np.array_equal(cm.expand_on_host().numpy(), want), k
        if isinstance(v, dict):
            assert np.array_equal(v["instance"].numpy(), z[f"c{case}.{k}.instance"])


def test_collate_and_byte_budget(gold):
    _, feats = gold
    samples = [data.compact_label_process(dict(label_case(c)[0]), label_case(c)[1], feats) for c in range(3)]
    batch = data.collate_experts(samples)
    assert batch["seg_coco"].u8.shape == (3, 1, 24, 24) and batch["seg_coco"].table.shape == (3, 256, 64)
    assert batch["obj_detection"]["label"].shape == (3, 64, 24, 24) and batch["obj_detection"]["instance"].shape == (3, 1, 24, 24)
    full = torch.stack([s["seg_coco"].expand_on_host() for s in samples])
    assert torch.equal(batch["seg_coco"].expand_on_host(), full)
    # at the reference's 224 x 224 label size: 50 KB map + 64 KB table instead of 12.8 MB per modality and image
    compact_bytes = 224 * 224 + 256 * 64 * 4
    assert compact_bytes * 100 < 64 * 224 * 224 * 4
    with pytest.raises(RuntimeError):
        batch["seg_coco"].expand()                                       # the model-side expansion is CUDA-only: no CPU fallback
